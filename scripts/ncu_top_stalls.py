"""Top stalled instructions per kernel from `ncu -i X.ncu-rep --page source --csv --print-kernel-base function`."""
import csv
import sys

path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
print(len(secs), "kernel sections")
s = secs[which]
hdr, body = s["rows"][0], [r for r in s["rows"][1:] if len(r) > 5]
c, si, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(float(r[c] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(s["name"], "total samples", tot)
agg = {st: sum(float(r[hdr.index(st)] or 0) for r in body) for st in stalls}
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(body, key=lambda r: -float(r[c] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    st = sorted(((float(r[hdr.index(x)] or 0), x) for x in stalls), reverse=True)[:2]
    print(f"{float(r[c]):7.0f} {100 * float(r[c]) / tot:5.1f}% ex={r[ie]:>8s} {r[si].strip()[:64]:64s} {st}")
