#!/bin/bash
# fp16 residual stream of the token layer: parity, timing against the fp32 stream, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "text or e2e or configs0 or cfg1 or coarse" 2>&1 | tail -n 25
timeout 300 python scripts/time_text.py 2>&1 | tail -n 2
T2L_TEXT_STREAM32=1 timeout 300 python scripts/time_text.py 2>&1 | tail -n 2
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_text_s16.csv python scripts/profile_step.py --skip-cells --cells 64 > gpurun_out/profile_launches.log 2>&1; echo "launch list text rc=$?"
python scripts/launch_summary.py gpurun_out/launches_text_s16.csv | head -n 14
