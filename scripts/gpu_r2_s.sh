#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -n 3
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "kernel tests failed rc=$rc"; timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" -s 2>&1 | tail -n 30; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "fine" 2>&1 | grep -E "passed|failed|error|offset" | head
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_fine_match.csv python scripts/profile_fine.py 3276 > $O/profile_fine.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py $O/launches_fine_match.csv > $O/launches_fine_match_summary.txt; head -n 9 $O/launches_fine_match_summary.txt
timeout 600 python bench.py --workload fine > $O/bench_fine.json 2> $O/bench_fine.err; echo "bench fine rc=$?"
python -c "
import json; d=json.loads(open('$O/bench_fine.json').read()); print(d['value'], d['detail'])"
