#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -s > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
grep -E "passed|failed" gpurun_out/t_kernels.log gpurun_out/t_parity.log
grep -E "differ|vs the reference|16 objects|rror|second pass|fallbacks" gpurun_out/t_parity.log | tail -n 20
timeout 300 python scripts/chunk_sweep.py 2>&1 | tail -n 1
timeout 300 python scripts/time_search.py 2>&1 | tail -n 4
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells_v4.csv python scripts/profile_step.py --cells 512 --queries 8 > gpurun_out/profile_launches.log 2>&1; echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/launches_cells_v4.csv 2>/dev/null | head -n 30
