#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search or merge or dropin" > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -n 3 gpurun_out/t_parity.log
./scripts/gpu_bench.sh
