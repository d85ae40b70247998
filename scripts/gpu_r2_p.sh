#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s > $O/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 1 $O/t_kernels.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > $O/t_parity.log 2>&1; echo "parity rc=$?"; tail -n 1 $O/t_parity.log
grep -E "error vs|embedding|offset|differ" $O/t_parity.log | head -n 12
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/smoke.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_fine_match.csv python scripts/profile_fine.py 3276 > $O/profile_fine.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py $O/launches_fine_match.csv > $O/launches_fine_match_summary.txt; head -n 9 $O/launches_fine_match_summary.txt
timeout 600 python bench.py --workload fine > $O/bench_fine.json 2> $O/bench_fine.err; echo "bench fine rc=$?"
python -c "
import json; d=json.loads(open('$O/bench_fine.json').read()); print(d['value'], d['detail'])"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_n1.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','cold_db_qps','gpu_launches','topk_matches_fp64_oracle_sample')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'enc frac', d['roofline_other_kernels']['db_encode']['frac'])
PY
