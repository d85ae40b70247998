#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -n 2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "text_golden or cell_golden or fine" 2>&1 | tail -n 2
timeout 200 python scripts/time_text.py 2>&1 | tail -n 1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_text_x.csv python scripts/profile_step.py --skip-cells --cells 64 > $O/profile_launches.log 2>&1; echo "launch list text rc=$?"
python scripts/launch_summary.py $O/launches_text_x.csv | head -n 6
timeout 300 python scripts/chunk_sweep.py 2>&1 | tail -n 1
