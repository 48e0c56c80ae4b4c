#!/usr/bin/env python
"""Builds one config's matrix on the device and launches one SpMV kernel variant a few times
(target of ncu captures).  Usage: python profiles/prof_one.py c3|c4|c2|c1 auto|stream|merge|split|vectorN [launches]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

cfg, kern = sys.argv[1], sys.argv[2]
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 5
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
if cfg == "c3":
    n = 10_000_000
    r, c, v = sd.random_uniform_coo_device(torch, n, 16, 8_000_000, torch.float32, seed=1)
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    tdt = torch.float32
elif cfg == "c4":
    scale = int(os.environ.get("RMAT_SCALE", "24"))
    n = 1 << scale
    r, c, v = sd.rmat_coo_device(torch, scale, 32, torch.float32, seed=3)
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    tdt = torch.float32
elif cfg == "c2":
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
    n, ptr, col, val = sd.stencil_device(torch, offs, 128, 26.0, -1.0, torch.float64)
    A = sp.CsrMatrix.from_device_arrays(n, n, col.numel(), ptr.data_ptr(), col.data_ptr(), val.data_ptr(), np.float64)
    tdt = torch.float64
else:
    offs = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]
    n, ptr, col, val = sd.stencil_device(torch, offs, 1024, 4.0, -1.0, torch.float64)
    A = sp.CsrMatrix.from_device_arrays(n, n, col.numel(), ptr.data_ptr(), col.data_ptr(), val.data_ptr(), np.float64)
    tdt = torch.float64
x = torch.rand(n, device="cuda", dtype=tdt) - 0.5
y = torch.empty(n, device="cuda", dtype=tdt)
k, l = (0, 0) if kern == "auto" else (5, 0) if kern == "stream" else (2, 0) if kern == "merge" else (3, 0) if kern == "split" else (1, int(kern[6:]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
A.spmv_device(x.data_ptr(), y.data_ptr(), k, l)
e0.record()
for _ in range(launches):
    A.spmv_device(x.data_ptr(), y.data_ptr(), k, l)
e1.record()
torch.cuda.synchronize()
print(cfg, kern, "nnz", A.nnz(), "ms/launch", e0.elapsed_time(e1) / launches)
