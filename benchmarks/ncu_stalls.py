#!/usr/bin/env python
"""Top stall locations (SASS lines) of the first kernel in a .ncu-rep captured with --import-source on."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
iS, iW = hdr.index("Source"), hdr.index("# Samples")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    try:
        n = int(r[iW])
    except Exception:
        continue
    data.append((n, r[iS].strip()[:84], {hdr[i]: int(r[i]) for i in cols if r[i] not in ("0", "")}))
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
agg = {}
for n, _, st in data:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print("by reason:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for n, src, st in sorted(data, key=lambda d: -d[0])[:top]:
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{n:7d} {100 * n / tot:5.1f}%  {src:84s} {best}")
