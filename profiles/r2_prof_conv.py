#!/usr/bin/env python
"""One call of each conversion / assembly on configs 1-3 (after a warm-up call): the target of an ncu launch
list (`--metrics gpu__time_duration.sum`), which gives the per-kernel times behind the round-2 work on
assembly and transposes.  Usage: python profiles/r2_prof_conv.py [c1 c2 c3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
which = sys.argv[1:] or ["c1", "c2", "c3"]


def timed(label, fn, reps=2):
    for i in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); out = fn(); b.record()
        torch.cuda.synchronize()
    print(f"{label}: {a.elapsed_time(b):.3f} ms", flush=True)
    return out


def shuffled(r, c, v, seed=42):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    p = torch.randperm(r.numel(), device="cuda", generator=g)
    return r[p].contiguous(), c[p].contiguous(), v[p].contiguous()


if "c1" in which:
    n, ptr, col, val = sd.stencil_device(torch, [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)], 1024, 4.0, -1.0, torch.float64)
    rows = torch.repeat_interleave(torch.arange(n, device="cuda", dtype=torch.int32), (ptr[1:] - ptr[:-1]).long())
    r, c, v = shuffled(rows, col, val)
    A = timed("c1 assembly shuffled", lambda: sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64))
    timed("c1 csr->csc", lambda: A.to_csc())
    del A, r, c, v, rows
if "c2" in which:
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
    n, ptr, col, val = sd.stencil_device(torch, offs, 128, 26.0, -1.0, torch.float64)
    A = sp.CsrMatrix.from_device_arrays(n, n, col.numel(), ptr.data_ptr(), col.data_ptr(), val.data_ptr(), np.float64, validate=False)
    timed("c2 csr->csc", lambda: A.to_csc())
    rows = torch.repeat_interleave(torch.arange(n, device="cuda", dtype=torch.int32), (ptr[1:] - ptr[:-1]).long())
    r, c, v = shuffled(rows, col, val)
    timed("c2 assembly shuffled", lambda: sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64))
    del A, r, c, v, rows, ptr, col, val
    torch.cuda.empty_cache()
if "c3" in which:
    n = 10_000_000
    r, c, v = sd.random_uniform_coo_device(torch, n, 16, 8_000_000, torch.float32, seed=1)
    A = timed("c3 assembly", lambda: sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32))
    del r, c, v
    timed("c3 csr->csc", lambda: A.to_csc())
