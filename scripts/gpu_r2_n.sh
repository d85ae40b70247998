#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -n 2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "fine or cell or text_golden" 2>&1 | grep -E "passed|failed|error|offset|embedding" | head
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02/launches_fine_match.csv python scripts/profile_fine.py 3277 > gpurun_out/r02/profile_fine.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/r02/launches_fine_match.csv | head -n 8
timeout 600 python bench.py --workload fine > gpurun_out/r02/bench_fine.json 2> gpurun_out/r02/bench_fine.err; echo "bench fine rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_fine.json').read()); print(d['value'], d['detail'])"
