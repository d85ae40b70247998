#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "text or e2e or host or drop" 2>&1 | tail -n 2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 900 python bench.py --steps 10 --warmup 3 --skip-extras > gpurun_out/r02/bench_n1_check.json 2> gpurun_out/r02/bench_n1_check.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_n1_check.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','gpu_launches','topk_matches_fp64_oracle_sample')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['clocks'])
PY
