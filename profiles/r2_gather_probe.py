"""One GPU, a world of one: the compute side of the fused gather kernel (tile producer + consumers) on a matrix shaped
like ONE block of config 3 at 8 GPUs (1.25 M rows, `per_row` random columns in a 1.25 M range, f32), against the
unsharded kernels on the same matrix.  Prints ms per product (back to back, CUDA events) and the kernel's own stamps.
    python profiles/r2_gather_probe.py"""
import json
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
os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
out = {}
n = 1_250_000
for per_row in (2, 4, 16):
    g = torch.Generator(device="cuda").manual_seed(per_row)
    rows = torch.arange(n, device="cuda", dtype=torch.int32).repeat_interleave(per_row)
    cols = torch.randint(0, n, (n * per_row,), device="cuda", generator=g, dtype=torch.int32)
    vals = torch.rand(n * per_row, device="cuda", generator=g, dtype=torch.float32) - 0.5
    D = spd.DistCsrMatrix.from_device_triplets(dist, torch, n, n, rows, cols, vals, ctx=ctx)
    nnz = D.local.nnz()
    x = torch.rand(n, device="cuda", dtype=torch.float32) - 0.5
    xv = spd.PeerVector(ctx, dist, n, np.float32, D.starts)
    device_view(torch, xv.local_ptr, n, torch.float32).copy_(x)
    xv.publish()
    xf = torch.empty(n, device="cuda", dtype=torch.float32)
    y = torch.empty(n, device="cuda", dtype=torch.float32)
    y0 = torch.empty(n, device="cuda", dtype=torch.float32)

    def timed(fn, batch=20, reps=5):
        for _ in range(3):
            fn()
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(batch):
                fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / batch)
        return best
    res = {"nnz": nnz}
    res["spl_spmv_ms"] = timed(lambda: D.local.spmv_device(x.data_ptr(), y0.data_ptr()))
    res["spl_spmv_choice"] = D.local.spmv_choice() if hasattr(D.local, "spmv_choice") else None
    res["spl_spmv_x_in_peer_buffer_ms"] = timed(lambda: D.local.spmv_device(xv.published_ptr, y0.data_ptr()))
    for lanes in (None, "1", "2", "4"):
        for per_sm in (None, "3", "2"):
            if lanes:
                os.environ["SPL_GATHER_LANES"] = lanes
            if per_sm:
                os.environ["SPL_GATHER_CTAS_PER_SM"] = per_sm
            try:
                D.prepare_gather(torch)
                ms = timed(lambda: D.spmv_gather(xv, xf.data_ptr(), y.data_ptr()))
                err = float(((y - y0).abs().max() / (y0.abs().max() + 1e-30)).item())
                res[f"fused lanes={lanes} ctas_per_sm={per_sm}"] = {"ms": ms, "rel_diff": err}
            except Exception as e:                                 # noqa: BLE001
                res[f"fused lanes={lanes} ctas_per_sm={per_sm}"] = str(e)[:80]
            os.environ.pop("SPL_GATHER_LANES", None)
            os.environ.pop("SPL_GATHER_CTAS_PER_SM", None)
    D.prepare_gather(torch)
    tl = torch.zeros(64, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    D.spmv_gather(xv, xf.data_ptr(), y.data_ptr(), timeline_dev=tl.data_ptr())
    torch.cuda.synchronize()
    t = tl.cpu().tolist()
    t0 = t[1]
    res["stamps_us"] = {"block": [(v - t0) / 1e3 for v in t[1:4]], "last_cta_done": (t[4] - t0) / 1e3}
    out[f"{per_row}_per_row"] = res
    xv.close(dist)
    del D
print(json.dumps(out, indent=1))
