#!/usr/bin/env python3
"""Summarise an ncu --page source --csv export: hottest SASS lines by stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = None
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        h = i
        break
hdr = rows[h]
iS, iI, iSrc, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Avg. Threads Executed")
stall_cols = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[h + 1:]:
    try:
        data.append((int(r[iS]), int(r[iI]), r[iSrc].strip(), r[iT], {hdr[c]: int(r[c]) for c in stall_cols}))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
toti = sum(d[1] for d in data)
print(f"total samples {tot}, warp instructions {toti}, SASS lines {len(data)}")
agg = {}
for d in data:
    for k, v in d[4].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for k, (s, i, src, thr_avg, st) in enumerate(data):
    if s > tot * thr / 100:
        top = max(st.items(), key=lambda x: x[1])
        print(f"{k:4d} {100*s/tot:5.1f}% smp  {100*i/toti:5.1f}% inst  thr {thr_avg:>5}  {top[0][6:]:>10}  {src[:90]}")
