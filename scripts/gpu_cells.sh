#!/bin/bash
# cell-encoder pass: kernel tests, encode parity, fused-vs-unfused A/B + timing, launch list of one encode chunk
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "fps or features2 or segmax" > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "encode_cells or dropin" > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
timeout 300 python scripts/ab_fused.py > gpurun_out/ab_fused.log 2>&1; echo "ab rc=$?"

timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cells.csv python scripts/profile_step.py --cells 512 --queries 8 > gpurun_out/profile_launches.log 2>&1; echo "rc=$?"
tail -n 4 gpurun_out/t_kernels.log gpurun_out/t_parity.log; cat gpurun_out/ab_fused.log | tail -5
python scripts/launch_summary.py gpurun_out/launches_cells.csv | head -14
