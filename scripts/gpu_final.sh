#!/bin/bash
# end-of-round evidence pass: all GPU tests, smoke, bench (engine + reference arm), launch lists, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
kill $SMI
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_text.csv python scripts/profile_step.py --skip-cells --cells 64 > gpurun_out/profile_launches.log 2>&1; echo "launch list text rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells.csv python scripts/profile_step.py --cells 512 --queries 8 >> gpurun_out/profile_launches.log 2>&1; echo "launch list cells rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:sa_obj -c 3 -o /tmp/prof_sa python scripts/profile_step.py --cells 512 --queries 8 > gpurun_out/prof_sa.log 2>&1; echo "ncu sa rc=$?"
ncu -i /tmp/prof_sa.ncu-rep --page raw --csv > gpurun_out/prof_sa_obj_raw.csv 2>/dev/null
tail -n 4 gpurun_out/t_kernels.log gpurun_out/t_parity.log gpurun_out/smoke.log; tail -c 600 gpurun_out/bench.err; cat gpurun_out/bench_ref.log | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','cold_db_qps','gpu_launches')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'enc frac', d['roofline_other_kernels']['db_encode']['frac'])
PY
