#!/bin/bash
# round 2, pass B: per-point/per-centroid first Linear in sa_obj2, fp16 single-pass search
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "features2 or fps or sa_empty" > gpurun_out/t_sa.log 2>&1; echo "sa tests rc=$?"
grep -E "error|passed|failed|Error" gpurun_out/t_sa.log | tail -n 8
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
grep -E "passed|failed|differ|vs the reference|16 objects|rror" gpurun_out/t_parity.log | tail -n 30
timeout 300 python scripts/chunk_sweep.py 2>&1 | tail -n 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells_v3.csv python scripts/profile_step.py --cells 512 --queries 8 > gpurun_out/profile_launches.log 2>&1; echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/launches_cells_v3.csv 2>/dev/null | head -n 16
timeout 300 python scripts/profile_search.py 2>&1 | tail -n 6
