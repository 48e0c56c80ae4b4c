"""One GPU, a world of one: a few launches of the fused gather kernel and of the stream kernel on the same matrix
(1.25 M rows, 2 random columns per row in a 1.25 M range, f32 — one block of config 3 at 8 GPUs), for ncu."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spalinalg_b200 as sp                                       # noqa: E402
from spalinalg_b200 import dist as spd                            # noqa: E402
from spalinalg_b200.synthetic_device import device_view           # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29534")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
n, per_row = 1_250_000, int(os.environ.get("PER_ROW", "2"))
g = torch.Generator(device="cuda").manual_seed(per_row)
rows = torch.arange(n, device="cuda", dtype=torch.int32).repeat_interleave(per_row)
cols = torch.randint(0, n, (n * per_row,), device="cuda", generator=g, dtype=torch.int32)
vals = torch.rand(n * per_row, device="cuda", generator=g, dtype=torch.float32) - 0.5
D = spd.DistCsrMatrix.from_device_triplets(dist, torch, n, n, rows, cols, vals, ctx=ctx)
x = torch.rand(n, device="cuda", dtype=torch.float32) - 0.5
xv = spd.PeerVector(ctx, dist, n, np.float32, D.starts)
device_view(torch, xv.local_ptr, n, torch.float32).copy_(x)
xv.publish()
xf = torch.empty(n, device="cuda", dtype=torch.float32)
y = torch.empty(n, device="cuda", dtype=torch.float32)
D.prepare_gather(torch)
for _ in range(3):
    D.local.spmv_device(x.data_ptr(), y.data_ptr())
    D.spmv_gather(xv, xf.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
print("ok")
xv.close(dist)
dist.destroy_process_group()
