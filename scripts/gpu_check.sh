#!/bin/bash
# regression pass: all GPU tests (one process per file), smoke, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 6 gpurun_out/t_kernels.log gpurun_out/t_parity.log gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench.err; cat gpurun_out/bench.log
