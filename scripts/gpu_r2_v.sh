#!/bin/bash
# DRAM traffic (ncu, metrics only) of every launch of one query step and of one 16 384-object encode chunk
O=gpurun_out/r02
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file $O/traffic_text_step.csv python scripts/profile_step.py --skip-cells --cells 64 > $O/traffic_text.log 2>&1; echo "text rc=$?"
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file $O/traffic_cells_chunk.csv python scripts/profile_step.py --cells 2048 --queries 8 > $O/traffic_cells.log 2>&1; echo "cells rc=$?"
tail -n 2 $O/traffic_text.log $O/traffic_cells.log
wc -l $O/traffic_text_step.csv $O/traffic_cells_chunk.csv
