#!/bin/bash
# ncu --set full + source-level stall sampling of the three sa_obj2 launches (4 096 objects)
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:${KPAT:-sa_obj2} -c 3 -o /tmp/prof_sa python scripts/profile_step.py --cells 512 --queries 8 > gpurun_out/prof_sa.log 2>&1; echo "capture rc=$?"
ncu -i /tmp/prof_sa.ncu-rep --page raw --csv > gpurun_out/prof_sa2_raw.csv 2>/dev/null
ncu -i /tmp/prof_sa.ncu-rep --page source --csv --print-kernel-base function > /tmp/prof_sa_src.csv 2>/dev/null
gzip -c /tmp/prof_sa_src.csv > gpurun_out/prof_sa2_src.csv.gz; ls -la gpurun_out/prof_sa2_src.csv.gz
for i in 0 1 2; do python scripts/ncu_top_stalls.py /tmp/prof_sa_src.csv $i 14; done
