#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_text.csv python scripts/profile_step.py --skip-cells --cells 64 > gpurun_out/profile_launches.log 2>&1; echo "rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells.csv python scripts/profile_step.py --cells 256 --queries 8 >> gpurun_out/profile_launches.log 2>&1; echo "rc=$?"
