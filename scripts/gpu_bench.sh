#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','cold_db_qps','gpu_launches')}, d['e2e']['value'])
print(json.dumps(d['roofline_other_kernels'], indent=1)[:1800])
PY
