#!/bin/bash
# ncu --set full on the token-layer GEMMs of one text chunk: launch order per chunk is
# QKV (0), out-proj (1), FFN1 (2), FFN2 (3) among the umma_gemm_kernel launches.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:umma_gemm_kernel --launch-skip 1 --launch-count 3 -o /tmp/prof_t python scripts/profile_step.py --skip-cells --cells 64 --queries 455 > gpurun_out/prof_t.log 2>&1; echo "capture rc=$?"
ncu -i /tmp/prof_t.ncu-rep --page raw --csv > gpurun_out/prof_t_raw.csv 2>/dev/null
ncu -i /tmp/prof_t.ncu-rep --page source --csv --print-kernel-base function > /tmp/prof_t_src.csv 2>/dev/null
for i in 0 1 2; do python scripts/ncu_top_stalls.py /tmp/prof_t_src.csv $((2*i)) 40 > gpurun_out/prof_t_stalls_$i.txt 2>&1; done
ls -la /tmp/prof_t.ncu-rep; head -c 600 gpurun_out/prof_t_stalls_0.txt
