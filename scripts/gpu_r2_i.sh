#!/bin/bash
# DB-encode round: new intra-cell attention core, qx1 rewrite, sa_obj2 gather trims
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -n 6
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "kernel tests failed rc=$rc"; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "cell or e2e or cfg1 or configs0 or stream" 2>&1 | grep -E "passed|failed|error|embedding|differ" | head -n 20
timeout 300 python scripts/chunk_sweep.py 2>&1 | tail -n 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells_16k.csv python scripts/profile_step.py --cells 2048 --queries 8 > gpurun_out/profile_launches.log 2>&1; echo "launch list cells rc=$?"
python scripts/launch_summary.py gpurun_out/launches_cells_16k.csv > gpurun_out/launches_cells_16k.txt; head -n 34 gpurun_out/launches_cells_16k.txt
