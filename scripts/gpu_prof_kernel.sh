#!/bin/bash
# ncu --set full on kernels matching $1 (regex) in the small profile run; raw csv + top stalls come back
PAT=${1:-sa_fused}
CNT=${2:-3}
ARGS=${3:---cells 256 --queries 8}
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:$PAT -c $CNT -o /tmp/prof_k python scripts/profile_step.py $ARGS > gpurun_out/prof_k.log 2>&1; echo "capture rc=$?"
ncu -i /tmp/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page source --csv --print-kernel-base function > /tmp/prof_k_src.csv 2>/dev/null
for i in $(seq 0 $((CNT-1))); do python scripts/ncu_top_stalls.py /tmp/prof_k_src.csv $((2*i)) 28 > gpurun_out/prof_k_stalls_$i.txt 2>&1; done
ls -la /tmp/prof_k.ncu-rep
