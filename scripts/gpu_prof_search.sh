#!/bin/bash
mkdir -p gpurun_out
T2L_SEARCH_FIRST=${1:-fp16} timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:umma_gemm_kernel -c 1 -o /tmp/prof_search python scripts/profile_search.py ${2:-100000} > gpurun_out/prof_search.log 2>&1; echo "capture rc=$?"
ncu -i /tmp/prof_search.ncu-rep --page raw --csv > gpurun_out/prof_search_raw.csv 2>/dev/null
ncu -i /tmp/prof_search.ncu-rep --page source --csv --print-kernel-base function > /tmp/prof_search_src.csv 2>/dev/null
gzip -c /tmp/prof_search_src.csv > gpurun_out/prof_search_src.csv.gz
python scripts/ncu_top_stalls.py /tmp/prof_search_src.csv 0 30
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_search_raw.csv')))
h=rows[0]; r=rows[2] if len(rows)>2 else rows[1]
for k in ('gpu__time_duration.sum','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_active','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.avg.per_cycle_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','dram__bytes_read.sum'):
    for i,c in enumerate(h):
        if c==k: print(k, r[i])
PY
