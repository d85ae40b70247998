#!/bin/bash
mkdir -p gpurun_out
MASKS="${MASKS:-0 1 2 6 8 14 15 16 32 47}"
timeout 900 ncu --profile-from-start off -k regex:${KPAT:-sa_fused} --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/sa_bisect.csv python scripts/sa_bisect.py $MASKS > gpurun_out/sa_bisect.log 2>&1; echo "rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/sa_bisect.csv')))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr=rows[hi]; mv=hdr.index('Metric Value')
us=[float(r[mv].replace(',',''))/1e3 for r in rows[hi+1:]]
masks="$MASKS".split()
print("mask   SA1    SA2    SA3  (us)")
for i,m in enumerate(masks):
    print(f"{m:>4s} "+" ".join(f"{x:6.0f}" for x in us[3*i:3*i+3]))
PY
