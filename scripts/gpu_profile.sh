#!/bin/bash
# ncu evidence: (1) per-launch durations of one encode + one query step, (2) --set full on the GEMMs
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_full.csv python scripts/profile_step.py > gpurun_out/profile_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:umma_gemm -c 14 -o /tmp/prof_umma python scripts/profile_step.py --small > gpurun_out/profile_full.log 2>&1; echo "full capture rc=$?"
ncu -i /tmp/prof_umma.ncu-rep --page raw --csv > gpurun_out/prof_umma_raw.csv 2>/dev/null
ls -la /tmp/prof_umma.ncu-rep
sz=$(stat -c %s /tmp/prof_umma.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_umma.ncu-rep gpurun_out/; fi
timeout 600 ncu --profile-from-start off --set full --clock-control none -f \
    -k regex:"edge_gather|mha_small|layer_norm|linear_simt|fps_kernel|ball_kernel|max_over_rows|rerank" -c 24 -o /tmp/prof_other python scripts/profile_step.py --small >> gpurun_out/profile_full.log 2>&1; echo "other capture rc=$?"
ncu -i /tmp/prof_other.ncu-rep --page raw --csv > gpurun_out/prof_other_raw.csv 2>/dev/null
du -sh gpurun_out
