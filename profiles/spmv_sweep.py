#!/usr/bin/env python
"""SpMV-only timing of the five configs (device-resident, CUDA events, median of 15 single launches,
L2 flushed for config 1).  Usage: python profiles/spmv_sweep.py [c1 c2 ...] [--kernels auto,split,vector4]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import run_configs as rc                                           # noqa: E402  (sets up ctx/stream)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

KERN = {"auto": (0, 0), "merge": (2, 0), "split": (3, 0)}
KERN.update({f"vector{l}": (1, l) for l in (1, 2, 4, 8, 16, 32)})


def build(cfg):
    t64, t32 = torch.float64, torch.float32
    if cfg == "c1":
        n, p, c, v = sd.stencil_device(torch, [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)], 1024, 4.0, -1.0, t64)
        return sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64, ctx=rc.ctx), 8, t64, True
    if cfg == "c2":
        offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
        n, p, c, v = sd.stencil_device(torch, offs, 128, 26.0, -1.0, t64)
        return sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64, ctx=rc.ctx), 8, t64, False
    if cfg == "c3":
        n = 10_000_000
        r, c, v = sd.random_uniform_coo_device(torch, n, 16, 8_000_000, t32, seed=1)
        return sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32, ctx=rc.ctx), 4, t32, False
    if cfg == "c4":
        n = 1 << 24
        r, c, v = sd.rmat_coo_device(torch, 24, 32, t32, seed=3)
        return sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32, ctx=rc.ctx), 4, t32, False
    n = 100_000_000
    p, c, v = sd.banded_device(torch, n, 0, n, range(-4, 5), t64)
    return sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64, ctx=rc.ctx), 8, t64, False


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    kernels = ["auto"]
    for a in sys.argv[1:]:
        if a.startswith("--kernels"):
            kernels = a.split("=", 1)[1].split(",")
    for cfg in args or ["c1", "c2", "c3", "c4", "c5"]:
        A, V, tdt, small = build(cfg)
        n, m, nnz = A.nrows(), A.ncols(), A.nnz()
        x = torch.rand(m, device="cuda", dtype=tdt) - 0.5
        y = torch.empty(n, device="cuda", dtype=tdt)
        b = nnz * (4 + V) + (n + m) * V
        row = {}
        for name in kernels:
            k, l = KERN[name]
            med, mn = rc.timed(lambda: A.spmv_device(x.data_ptr(), y.data_ptr(), k, l), reps=15, warm=3, flush_l2=small)
            row[name] = (round(med, 4), round(b / med / 1e6 / rc.PEAK, 3))
        print(cfg, "planned", A.spmv_choice(), row, flush=True)
        del A, x, y
        torch.cuda.empty_cache()
