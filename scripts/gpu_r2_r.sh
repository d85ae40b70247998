#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
TOOLS="memcheck synccheck racecheck" PARTS="fine" bash scripts/gpu_sanitize.sh 2>&1 | tee $O/sanitize_summary_fine.txt
cp gpurun_out/sanitize_*_fine.log $O/ 2>/dev/null
timeout 600 ncu --profile-from-start off --set full --clock-control none -f \
    -k regex:mha_seq_kernel -c 1 -o /tmp/prof_mha python scripts/profile_step.py --cells 2048 --queries 8 > $O/prof_mha.log 2>&1; echo "ncu mha rc=$?"
ncu -i /tmp/prof_mha.ncu-rep --page raw --csv > $O/ncu_full_mha_seq_raw.csv 2>/dev/null
ls -la $O | head
