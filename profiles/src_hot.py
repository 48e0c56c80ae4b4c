#!/usr/bin/env python
"""Hottest SASS instructions of an `ncu --page source --csv` dump: samples, executed count, stalls.
Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/src_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("kernel:", rows[0][1][:100], "| total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_")]
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
    s = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{100 * s / tot:5.1f}%  exec {r[col['Instructions Executed']]:>10}  thr/inst {r[col['Avg. Threads Executed']]:>5}  "
          f"{'/'.join(f'{h}:{c}' for c, h in why if c):32s} {r[col['Source']][:90]}")
