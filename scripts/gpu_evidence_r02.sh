#!/bin/bash
# round-2 evidence pass: sanitizers, launch lists, ncu --set full of the three named kernels
mkdir -p gpurun_out/r02
TOOLS="memcheck synccheck racecheck" PARTS="cells text search" bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r02/sanitize_summary.txt
cp gpurun_out/sanitize_*.log gpurun_out/r02/ 2>/dev/null
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02/launches_text.csv python scripts/profile_step.py --skip-cells --cells 64 > gpurun_out/r02/profile_launches.log 2>&1; echo "launch list text rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02/launches_cells.csv python scripts/profile_step.py --cells 512 --queries 8 >> gpurun_out/r02/profile_launches.log 2>&1; echo "launch list cells rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02/launches_search.csv python scripts/profile_search.py >> gpurun_out/r02/profile_launches.log 2>&1; echo "launch list search rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:sa_obj2 -c 3 -o /tmp/prof_sa python scripts/profile_step.py --cells 512 --queries 8 > gpurun_out/r02/prof_sa.log 2>&1; echo "ncu sa rc=$?"
ncu -i /tmp/prof_sa.ncu-rep --page raw --csv > gpurun_out/r02/ncu_full_sa_obj2_raw.csv 2>/dev/null
ncu -i /tmp/prof_sa.ncu-rep --page source --csv --print-kernel-base function > /tmp/prof_sa_src.csv 2>/dev/null
for i in 0 1 2; do python scripts/ncu_top_stalls.py /tmp/prof_sa_src.csv $i 12; done > gpurun_out/r02/ncu_stalls_sa_obj2.txt 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:umma_gemm_kernel -c 1 -o /tmp/prof_search python scripts/profile_search.py > gpurun_out/r02/prof_search.log 2>&1; echo "ncu search rc=$?"
ncu -i /tmp/prof_search.ncu-rep --page raw --csv > gpurun_out/r02/ncu_full_search_topk_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -f \
    -k regex:umma_gemm_kernel -s 0 -c 4 -o /tmp/prof_text python scripts/profile_step.py --skip-cells --cells 64 --queries 455 > gpurun_out/r02/prof_text.log 2>&1; echo "ncu text rc=$?"
ncu -i /tmp/prof_text.ncu-rep --page raw --csv > gpurun_out/r02/ncu_full_text_gemms_raw.csv 2>/dev/null
python scripts/launch_summary.py gpurun_out/r02/launches_cells.csv | head -n 30
python scripts/launch_summary.py gpurun_out/r02/launches_text.csv | head -n 16
ls -la gpurun_out/r02
