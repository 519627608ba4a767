#!/usr/bin/env python3
"""Per CUDA-source-line instruction / stall-sample breakdown of an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py X.ncu-rep [min_pct]   (all source files of the kernel, first profiled launch)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
data, cur, hdr, seen_kernel = [], None, None, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if r[0] == "Address" or hdr is None:
        hdr = None if r[0] == "Address" else hdr
        continue
    try:
        data.append((cur, int(r[0]), r[1], int(r[iS]), int(r[iI]), int(r[iT])))
    except (ValueError, IndexError):
        pass
tot = sum(d[4] for d in data)
tots = sum(d[3] for d in data)
print(f"total warp instructions {tot}, samples {tots}")
for d in data:
    if d[4] > tot * thr / 100 or d[3] > tots * thr / 100:
        print(f"{(d[0] or '?')[:14]:14s}{d[1]:5d} inst {100*d[4]/tot:5.1f}% smp {100*d[3]/max(tots,1):5.1f}% thr {d[5]/max(d[4],1):5.1f} | {d[2].strip()[:100]}")
