import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import spalinalg_b200 as sp
from spalinalg_b200 import synthetic_device as sd
n = int(float(sys.argv[1]))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream); sp.set_default_context(ctx)
r, c, v = sd.random_uniform_coo_device(torch, n, 16, n * 16 // 20, torch.float32, seed=1)
A = None
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    B = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    keep = (rep % 2 == 0)
    print(f"rep {rep} {1e3*(t1-t0):.2f} ms  (previous result alive: {A is not None})", flush=True)
    A = B if keep else None
    del B
