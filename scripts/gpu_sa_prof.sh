#!/bin/bash
# ncu --set full + source-level stall sampling of the three sa_fused kernels for one T2L_SA_DBG mask
MASK=${1:-0}
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:${KPAT:-sa_fused} -c 3 -o /tmp/prof_sa python scripts/sa_bisect.py $MASK > gpurun_out/prof_sa.log 2>&1; echo "capture rc=$?"
ncu -i /tmp/prof_sa.ncu-rep --page raw --csv > gpurun_out/prof_sa_raw_$MASK.csv 2>/dev/null
ncu -i /tmp/prof_sa.ncu-rep --page source --csv --print-source sass,cuda --print-kernel-base function > /tmp/prof_sa_src.csv 2>/dev/null || \
ncu -i /tmp/prof_sa.ncu-rep --page source --csv --print-kernel-base function > /tmp/prof_sa_src.csv 2>/dev/null
head -c 2000 /tmp/prof_sa_src.csv | head -5
gzip -c /tmp/prof_sa_src.csv > gpurun_out/prof_sa_src_$MASK.csv.gz; ls -la gpurun_out/prof_sa_src_$MASK.csv.gz
