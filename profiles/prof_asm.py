#!/usr/bin/env python
"""COO->CSR assembly of config-3 style random triplets (f32, 16/row + 5 % duplicates) at a chosen
row count; prints ms per assembly.  Target of launch lists / ncu captures.
Usage: python profiles/prof_asm.py [nrows] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
r, c, v = sd.random_uniform_coo_device(torch, n, 16, n * 16 // 20, torch.float32, seed=1)
for _ in range(3):      # the first calls grow the stream-ordered memory pool (hundreds of ms, once)
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"assembly n={n} len={r.numel()} nnz={A.nnz()} ms={ms:.3f} Mnnz/s={r.numel() / ms / 1e3:.0f}")
