#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -n 5 gpurun_out/t_kernels.log; grep -E "error|passed|failed|fallbacks" gpurun_out/t_parity.log | tail -n 20
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_text.csv python scripts/profile_step.py --skip-cells --cells 64 > gpurun_out/profile_launches.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_text.csv 0 10
./scripts/gpu_bench.sh 2>&1 | head -3
