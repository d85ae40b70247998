#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02/launches_fine_match.csv python scripts/profile_fine.py 3277 > gpurun_out/r02/profile_fine.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/r02/launches_fine_match.csv | head -n 30
