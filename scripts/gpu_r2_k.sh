#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -k regex:sa_obj2 \
    --log-file gpurun_out/sa2_bisect.csv python scripts/sa_bisect.py > gpurun_out/sa2_bisect.log 2>&1; echo "rc=$?"
grep gpu__time gpurun_out/sa2_bisect.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g'
