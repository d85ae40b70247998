#!/bin/bash
# TMA residual epilogue: kernel test first (bounded), then text parity, timing A/B, launch list
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "residual_stream or f16_operands" 2>&1 | tail -n 30
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "kernel test failed rc=$rc"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "text" 2>&1 | tail -n 12
timeout 200 python scripts/time_text.py 2>&1 | tail -n 1
T2L_TEXT_REG_EPI=1 timeout 200 python scripts/time_text.py 2>&1 | tail -n 1
T2L_TEXT_STREAM32=1 timeout 200 python scripts/time_text.py 2>&1 | tail -n 1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_text_tma.csv python scripts/profile_step.py --skip-cells --cells 64 > gpurun_out/profile_launches.log 2>&1; echo "launch list text rc=$?"
python scripts/launch_summary.py gpurun_out/launches_text_tma.csv > gpurun_out/launches_text_tma.txt; head -n 12 gpurun_out/launches_text_tma.txt
