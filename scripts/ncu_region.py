"""Print the SASS (address order) around the tensor-core issue loop with stall samples, from an ncu source-page csv."""
import csv
import sys

path, which = sys.argv[1], int(sys.argv[2])
pat = sys.argv[3] if len(sys.argv) > 3 else "UTCHMMA"
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
s = secs[which]
hdr, body = s["rows"][0], [r for r in s["rows"][1:] if len(r) > 5]
c, si, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = [i for i, r in enumerate(body) if pat in r[si]]
lo, hi = max(0, idx[0] - 45), min(len(body), idx[-1] + 30)
for r in body[lo:hi]:
    st = sorted(((float(r[hdr.index(x)] or 0), x.replace("stall_", "")) for x in stalls), reverse=True)[:2]
    st = [f"{n}:{int(v)}" for v, n in st if v > 0]
    print(f"{int(float(r[c] or 0)):6d} ex={r[ie]:>9s} {r[si].strip()[:70]:70s} {' '.join(st)}")
