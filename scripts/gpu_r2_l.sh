#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -n 3
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "kernel tests failed rc=$rc"; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "cell or e2e or cfg1 or configs0" 2>&1 | grep -E "passed|failed|error|embedding|differ" | head -n 20
timeout 300 python scripts/chunk_sweep.py 2>&1 | tail -n 1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -k regex:sa_obj2 \
    --log-file gpurun_out/sa2_bisect.csv python scripts/sa_bisect.py > gpurun_out/sa2_bisect.log 2>&1; echo "rc=$?"
grep gpu__time gpurun_out/sa2_bisect.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g'
