#!/usr/bin/env python
"""Hottest SASS lines of an `ncu --page source --csv` dump: samples, executed count, dominant stall.
    python tools/ncu_src_top.py prof_source.csv [N] [lo:hi]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
start = max(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name")
hdr = rows[start + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[start + 2:]
si, ei, sa = ix["Source"], ix["Instructions Executed"], ix["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_")]
f = lambda v: float(v) if v not in ("", None) else 0.0
tot_e = sum(f(r[ei]) for r in data)
tot_s = sum(f(r[sa]) for r in data)
print(f"{len(data)} SASS lines, {tot_e:.3e} warp instructions, {tot_s:.0f} samples")
if len(sys.argv) > 3:
    lo, hi = (int(v) for v in sys.argv[3].split(":"))
    for i in range(lo, hi):
        r = data[i]
        dom = max(stalls, key=lambda h: f(r[ix[h]]))
        print(f"{i:5d} {int(f(r[sa])):7d} {int(f(r[ei])):10d} {dom[6:]:18s} {r[si][:100]}")
    sys.exit()
for i, r in sorted(enumerate(data), key=lambda kv: -f(kv[1][sa]))[:n_top]:
    dom = max(stalls, key=lambda h: f(r[ix[h]]))
    print(f"{i:5d} {int(f(r[sa])):7d} {int(f(r[ei])):10d} {dom[6:]:18s} {r[si][:100]}")
