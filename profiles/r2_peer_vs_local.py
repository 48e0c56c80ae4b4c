#!/usr/bin/env python
"""One GPU, world of one: the same banded rows through spl_spmv (XLocal gather) and spl_spmv_peer (XPeer gather,
stream and vector kernels) - isolates what the owner test in the gather costs.  -> stdout"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import dist as spd                             # noqa: E402
from spalinalg_b200.synthetic_device import banded_device, device_view   # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
n = 50_000_000
p, c, v = banded_device(torch, n, 0, n, range(-4, 5), torch.float64)
A = sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64, validate=False, ctx=ctx)
del p, c, v
b = A.nnz() * 12 + 2 * n * 8
x = torch.sin(torch.arange(n, device="cuda", dtype=torch.float64) * 1e-3)
y = torch.empty(n, device="cuda", dtype=torch.float64)
xv = spd.PeerVector(ctx, dist, n, np.float64, [0, n])
for _ in range(2):
    device_view(torch, xv.local_ptr, n, torch.float64).copy_(x)
    xv.publish()
dA = spd.DistCsrMatrix(A, [0, n], 0, n, n)


def rate(fn, reps=20):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(3):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(e) / reps)
    return best


for name, fn, env in (("local stream", lambda: A.spmv_device(x.data_ptr(), y.data_ptr()), {}),
                      ("local vector", lambda: A.spmv_device(x.data_ptr(), y.data_ptr(), 1, 0), {}),
                      ("peer stream", lambda: dA.spmv_peer(xv, y.data_ptr()), {}),
                      ("peer stream + barrier", lambda: (xv.barrier(), dA.spmv_peer(xv, y.data_ptr())), {}),
                      ("peer vector", lambda: dA.spmv_peer(xv, y.data_ptr()), {"SPL_PEER_VECTOR": "1"})):
    for k, val in env.items():
        os.environ[k] = val
    ms = rate(fn)
    for k in env:
        os.environ.pop(k)
    print(f"{name:24s} {ms:.4f} ms  {b / ms / 1e6:8.1f} GB/s  frac {b / ms / 1e6 / 6547.5:.3f}", flush=True)
dist.destroy_process_group()
