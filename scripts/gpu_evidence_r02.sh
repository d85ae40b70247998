#!/bin/bash
# round-2 evidence pass: all GPU tests, smoke, bench (engine + reference arm + fine stage), sanitizers, launch lists,
# ncu --set full of the named kernels.  Everything lands in gpurun_out/r02/ and is copied to profiles/r02/ by hand.
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s > $O/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 1 $O/t_kernels.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > $O/t_parity.log 2>&1; echo "parity rc=$?"; tail -n 1 $O/t_parity.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload fine > $O/bench_fine.json 2> $O/bench_fine.err; echo "bench fine rc=$?"
TOOLS="memcheck synccheck racecheck" PARTS="cells text search" bash scripts/gpu_sanitize.sh 2>&1 | tee $O/sanitize_summary.txt
cp gpurun_out/sanitize_*.log $O/ 2>/dev/null
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_text.csv python scripts/profile_step.py --skip-cells --cells 64 > $O/profile_launches.log 2>&1; echo "launch list text rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_cells.csv python scripts/profile_step.py --cells 2048 --queries 8 >> $O/profile_launches.log 2>&1; echo "launch list cells rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_search.csv python scripts/profile_search.py >> $O/profile_launches.log 2>&1; echo "launch list search rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:sa_obj2 -c 3 -o /tmp/prof_sa python scripts/profile_step.py --cells 512 --queries 8 > $O/prof_sa.log 2>&1; echo "ncu sa rc=$?"
ncu -i /tmp/prof_sa.ncu-rep --page raw --csv > $O/ncu_full_sa_obj2_raw.csv 2>/dev/null
ncu -i /tmp/prof_sa.ncu-rep --page source --csv --print-kernel-base function > /tmp/prof_sa_src.csv 2>/dev/null
for i in 0 1 2; do python scripts/ncu_top_stalls.py /tmp/prof_sa_src.csv $i 12; done > $O/ncu_stalls_sa_obj2.txt 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:umma_gemm_kernel -c 1 -o /tmp/prof_search python scripts/profile_search.py > $O/prof_search.log 2>&1; echo "ncu search rc=$?"
ncu -i /tmp/prof_search.ncu-rep --page raw --csv > $O/ncu_full_search_topk_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -f \
    -k regex:umma_gemm_kernel -s 0 -c 4 -o /tmp/prof_text python scripts/profile_step.py --skip-cells --cells 64 --queries 455 > $O/prof_text.log 2>&1; echo "ncu text rc=$?"
ncu -i /tmp/prof_text.ncu-rep --page raw --csv > $O/ncu_full_text_gemms_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -f \
    -k regex:mha_seq_kernel -c 1 -o /tmp/prof_mha python scripts/profile_step.py --cells 2048 --queries 8 > $O/prof_mha.log 2>&1; echo "ncu mha rc=$?"
ncu -i /tmp/prof_mha.ncu-rep --page raw --csv > $O/ncu_full_mha_seq64_raw.csv 2>/dev/null
python scripts/launch_summary.py $O/launches_cells.csv > $O/launches_cells_summary.txt 2>&1
python scripts/launch_summary.py $O/launches_text.csv > $O/launches_text_summary.txt 2>&1
python scripts/launch_summary.py $O/launches_search.csv > $O/launches_search_summary.txt 2>&1
head -n 12 $O/launches_cells_summary.txt; head -n 10 $O/launches_text_summary.txt
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_n1.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','cold_db_qps','gpu_launches','topk_matches_fp64_oracle_sample')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'enc frac', d['roofline_other_kernels']['db_encode']['frac'])
for f in ('bench_ref_n1.json','bench_fine.json'):
    print(open('gpurun_out/r02/'+f).read()[:600])
PY
ls -la $O | head -50
