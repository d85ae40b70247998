#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "text or dropin" > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -n 12 gpurun_out/t_parity.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_text.csv python scripts/profile_step.py --skip-cells --cells 64 > gpurun_out/profile_launches.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_text.csv 0 30
