#!/usr/bin/env python
"""A * A on the 2-D Laplacian (config 1 matrix): ms per product on the device and, on a 256^2 sample,
the oracle port of the reference's Mul on one CPU core.  Usage: python profiles/prof_mul.py [grid] [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
if len(sys.argv) > 3 and sys.argv[3] == "3d":      # 27-point stencil g^3 (729 products per row)
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
    n, p, c, v = sd.stencil_device(torch, offs, g, 26.0, -1.0, torch.float64)
else:
    n, p, c, v = sd.stencil_device(torch, [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)], g, 4.0, -1.0, torch.float64)
A = sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64, ctx=ctx)
for _ in range(2):
    Cm = A * A
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    Cm = A * A
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"A*A stencil grid {g}: nnz(A)={A.nnz()} nnz(C)={Cm.nnz()} ms={ms:.3f}  ({Cm.nnz() / ms / 1e3:.0f} M entries of C per second)")
