#!/usr/bin/env python
"""Config-5 style sparse add on one GPU (A = band -4..4, B = offsets {-8,-2,0,2,8}); prints ms per
A + B.  Target of launch lists / ncu captures.  Usage: python profiles/prof_add.py [n] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
mats = []
for offs in (range(-4, 5), (-8, -2, 0, 2, 8)):
    p, c, v = sd.banded_device(torch, n, 0, n, offs, torch.float64)
    mats.append(sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(),
                                                np.float64, validate=False, ctx=ctx))
    del p, c, v
A, B = mats
Cm = A + B
del Cm
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    Cm = A + B
    del Cm
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nnz_c = (A + B).nnz()
b = (A.nnz() + B.nnz() + nnz_c) * 12 + 3 * (n + 1) * 4
print(f"add n={n} nnzA={A.nnz()} nnzB={B.nnz()} nnzC={nnz_c} ms={ms:.3f} algorithmic GB/s={b / ms / 1e6:.1f}")
