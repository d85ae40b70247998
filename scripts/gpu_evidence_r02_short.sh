#!/bin/bash
# trimmed evidence pass after a late kernel change: GPU tests, smoke, bench, sanitizers (cells, text), launch lists, text GEMM capture
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s > $O/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 1 $O/t_kernels.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > $O/t_parity.log 2>&1; echo "parity rc=$?"; tail -n 1 $O/t_parity.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
TOOLS="memcheck synccheck racecheck" PARTS="cells text" bash scripts/gpu_sanitize.sh 2>&1 | tee $O/sanitize_summary_cells_text.txt
cp gpurun_out/sanitize_*_cells.log gpurun_out/sanitize_*_text.log $O/ 2>/dev/null
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_text.csv python scripts/profile_step.py --skip-cells --cells 64 > $O/profile_launches.log 2>&1; echo "launch list text rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_cells.csv python scripts/profile_step.py --cells 2048 --queries 8 >> $O/profile_launches.log 2>&1; echo "launch list cells rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -f \
    -k regex:umma_gemm_kernel -s 0 -c 4 -o /tmp/prof_text python scripts/profile_step.py --skip-cells --cells 64 --queries 455 > $O/prof_text.log 2>&1; echo "ncu text rc=$?"
ncu -i /tmp/prof_text.ncu-rep --page raw --csv > $O/ncu_full_text_gemms_raw.csv 2>/dev/null
python scripts/launch_summary.py $O/launches_cells.csv > $O/launches_cells_summary.txt 2>&1
python scripts/launch_summary.py $O/launches_text.csv > $O/launches_text_summary.txt 2>&1
head -n 8 $O/launches_text_summary.txt
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_n1.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','cold_db_qps','gpu_launches','topk_matches_fp64_oracle_sample')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'enc frac', d['roofline_other_kernels']['db_encode']['frac'], d['roofline_other_kernels']['token_ffn1_alone']['achieved'], d['clocks'])
PY
