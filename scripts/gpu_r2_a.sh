#!/bin/bash
# round 2, pass A: correctness of sa_obj2 + new tests, A/B encode throughput, launch lists of both SA generations
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "features2 or fps or sa_empty" > gpurun_out/t_sa.log 2>&1; echo "sa tests rc=$?"
tail -n 12 gpurun_out/t_sa.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -s > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -n 15 gpurun_out/t_kernels.log; grep -E "passed|failed|differ|vs the reference|16 objects|Error|error" gpurun_out/t_parity.log | tail -n 30
timeout 300 python scripts/chunk_sweep.py 2>&1 | tail -n 1
T2L_SA_V1=1 timeout 300 python scripts/chunk_sweep.py 2>&1 | tail -n 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells_v2.csv python scripts/profile_step.py --cells 512 --queries 8 > gpurun_out/profile_launches.log 2>&1; echo "launch list v2 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_cells_v2.csv | head -n 24
T2L_SA_V1=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells_v1.csv python scripts/profile_step.py --cells 512 --queries 8 >> gpurun_out/profile_launches.log 2>&1; echo "launch list v1 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_cells_v1.csv | head -n 12
