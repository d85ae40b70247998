#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
timeout 300 python scripts/ab_fused.py > gpurun_out/ab_fused.log 2>&1; echo "ab rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/t_kernels.log gpurun_out/t_parity.log gpurun_out/ab_fused.log; tail -c 600 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_text_head','ms_search','db_encode_cells_per_s','cold_db_qps','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['cublas_tf32_same_shape_tflops'])
PY
