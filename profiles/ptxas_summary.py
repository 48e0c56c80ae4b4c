#!/usr/bin/env python
"""Compact per-kernel resource table from the nvcc -Xptxas -v logs under spalinalg_b200/build/.
Usage: python profiles/ptxas_summary.py [substring ...]"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pats = sys.argv[1:]
for log in sorted(glob.glob(os.path.join(ROOT, "spalinalg_b200", "build", "*.o.log"))):
    name = None
    spill = ""
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = f"stack {m.group(1)} spill {m.group(2)}/{m.group(3)}"
            continue
        m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?$", line.strip())
        if m and name:
            smem = re.search(r"(\d+) bytes smem", line)
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"\(anonymous namespace\)::|spl::|^void ", "", dem)
            dem = dem.split("(")[0]
            if not pats or any(p in dem for p in pats):
                print(f"{os.path.basename(log)[:-6]:12s} regs {m.group(1):>3s} smem {smem.group(1) if smem else '0':>6s} {spill:28s} {dem[:110]}")
            name, spill = None, ""
