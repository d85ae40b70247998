#!/bin/bash
# text-head pass: f16 GEMM test, text parity, bench (A/B vs tf32)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "f16 or tf32_emulation" > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "text or dropin" > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -n 25 gpurun_out/t_kernels.log; tail -n 12 gpurun_out/t_parity.log
./scripts/gpu_bench.sh
