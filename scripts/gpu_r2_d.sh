#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/time_search.py 2>&1 | tail -n 4
T2L_SEARCH_FIRST=fp16 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_search_fp16.csv python scripts/profile_search.py > gpurun_out/profile_search.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_search_fp16.csv 2>/dev/null | head -n 14
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k search 2>&1 | tail -n 2
bash scripts/gpu_prof_sa2.sh 2>&1 | tail -n 60
