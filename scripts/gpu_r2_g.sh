#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/tok_chunk_sweep.py 2>&1 | tail -n 8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tma.log 2> gpurun_out/bench_tma.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tma.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','cold_db_qps','gpu_launches','topk_matches_fp64_oracle_sample')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'enc frac', d['roofline_other_kernels']['db_encode']['frac'])
PY
