"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel, and optionally list launches in order."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
names = [re.sub(r"\(.*", "", r[kn]).replace("void t2l::", "").replace("t2l::", "") for r in data]
us = [float(r[mv].replace(",", "")) / 1e3 for r in data]
agg = collections.OrderedDict()
for n, t in zip(names, us):
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(us)
print(f"{len(data)} launches, {tot / 1e3:.3f} ms total")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {k[:100]}")
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for i in range(a, min(b, len(data))):
        print(f"{us[i]:9.1f} us  grid {data[i][gs]:>14s}  {names[i][:90]}")
