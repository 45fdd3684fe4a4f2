#!/usr/bin/env python
"""Top stall-sampled SASS lines of an `ncu --page source --csv` dump (second line is the header)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
col, src, ex = ci["Warp Stall Sampling (All Samples)"], ci["Source"], ci["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        v = float(r[col])
    except ValueError:
        continue
    top = sorted(((float(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
    data.append((v, r[src].strip()[:70], r[ex], ",".join(f"{s[6:]}:{int(x)}" for x, s in top if x > 0)))
tot = sum(d[0] for d in data)
print("total samples", tot)
for v, s, e, t in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v:8.0f} {100 * v / tot:5.1f}%  exec={e:>9}  {s:70s} {t}")
