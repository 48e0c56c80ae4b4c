#!/usr/bin/env python
"""Round 2 SpMV sweeps on one B200: the stream kernel (stages x CTAs per SM, with and without
programmatic dependent launch) against the vector kernel on configs 1, 2, 5 and 3, and the hot-column
nnz-split kernel (table size x threads per CTA) on config 4.  Back-to-back launches on one stream,
CUDA events around the batch; footprints below 2x L2 rotate over independent copies of (A, x, y).
Usage: python profiles/r2_spmv_sweep.py [c1 c2 c3 c4 c5]   -> gpurun_out/r2_spmv_sweep.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
OUT = {}


def rate(As, xs, ys, kernel, lanes=0, reps=200, warm=10):
    k = len(As)
    for i in range(warm):
        As[i % k].spmv_device(xs[i % k].data_ptr(), ys[i % k].data_ptr(), kernel, lanes)
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(reps):
            As[i % k].spmv_device(xs[i % k].data_ptr(), ys[i % k].data_ptr(), kernel, lanes)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def sweep_regular(name, make, V, tdt, copies, reps):
    As, xs, ys = [], [], []
    for _ in range(copies):
        A = make()
        As.append(A)
        xs.append(torch.rand(A.ncols(), device="cuda", dtype=tdt) - 0.5)
        ys.append(torch.empty(A.nrows(), device="cuda", dtype=tdt))
    A = As[0]
    b = A.nnz() * (4 + V) + (A.nrows() + A.ncols()) * V
    res = {"bytes": b, "planned": A.spmv_choice(), "copies": copies}

    def put(key, ms):
        res[key] = {"ms": round(ms, 5), "gbps": round(b / ms / 1e6, 1), "frac": round(b / ms / 1e6 / PEAK, 4)}
        print(name, key, res[key], flush=True)

    knobs = dict(SPL_STREAM_STAGES=None, SPL_STREAM_CTAS=None, SPL_NO_PDL=None, SPL_STREAM_TIGHT=None, SPL_STREAM_L2HINT=None)
    setenv(**knobs)
    put("vector", rate(As, xs, ys, 1, 0, reps))
    # check the stream kernel against the vector kernel once: same lanes, same order => same bits
    y_ref = ys[0].clone()
    As[0].spmv_device(xs[0].data_ptr(), ys[0].data_ptr(), 5, 0)
    torch.cuda.synchronize()
    res["stream_equals_vector_bitwise"] = bool(torch.equal(ys[0], y_ref))
    put("stream_default", rate(As, xs, ys, 5, 0, reps))
    setenv(SPL_NO_PDL=1)
    put("stream_default_nopdl", rate(As, xs, ys, 5, 0, reps))
    setenv(SPL_NO_PDL=None)
    for hint in (0, 1):
        for tight in (0, 1):
            for stages in (2, 3):
                setenv(SPL_STREAM_STAGES=stages, SPL_STREAM_TIGHT=tight, SPL_STREAM_L2HINT=hint)
                key = f"stream_hint{hint}_tight{tight}_s{stages}"
                try:
                    put(key, rate(As, xs, ys, 5, 0, reps))
                except Exception as e:                                   # noqa: BLE001
                    res[key] = {"error": str(e)[:60]}
                    print(name, key, "error", str(e)[:60], flush=True)
    setenv(**knobs)
    if copies > 1:
        put("vector_l2_resident", rate(As[:1], xs[:1], ys[:1], 1, 0, reps))
        put("stream_l2_resident", rate(As[:1], xs[:1], ys[:1], 5, 0, reps))
    OUT[name] = res


def c1():
    n, p, c, v = sd.stencil_device(torch, [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)], 1024, 4.0, -1.0, torch.float64)
    mk = lambda: sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64,
                                                 validate=False, ctx=ctx)
    sweep_regular("c1", mk, 8, torch.float64, 4, 400)


def c2():
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
    n, p, c, v = sd.stencil_device(torch, offs, 128, 26.0, -1.0, torch.float64)
    mk = lambda: sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64,
                                                 validate=False, ctx=ctx)
    sweep_regular("c2", mk, 8, torch.float64, 1, 100)


def c3():
    n = 10_000_000
    r, c, v = sd.random_uniform_coo_device(torch, n, 16, 8_000_000, torch.float32, seed=1)
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32, ctx=ctx)
    del r, c, v
    sweep_regular("c3", lambda: A, 4, torch.float32, 1, 30)


def c5():
    n = 100_000_000
    p, c, v = sd.banded_device(torch, n, 0, n, range(-4, 5), torch.float64)
    A = sp.CsrMatrix.from_device_arrays(n, n, c.numel(), p.data_ptr(), c.data_ptr(), v.data_ptr(), np.float64,
                                        validate=False, ctx=ctx)
    del p, c, v
    sweep_regular("c5", lambda: A, 8, torch.float64, 1, 20)


def c4():
    n = 1 << 24
    r, c, v = sd.rmat_coo_device(torch, 24, 32, torch.float32, seed=3)
    A0 = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32, ctx=ctx)
    del r, c, v
    torch.cuda.empty_cache()
    nnz = A0.nnz()
    b = nnz * 8 + 2 * n * 4
    res = {"bytes": b, "nnz": nnz}
    x = torch.rand(n, device="cuda", dtype=torch.float32) - 0.5
    y = torch.empty(n, device="cuda", dtype=torch.float32)
    pp, pi, pv = A0.device_ptrs()
    y_ref = None
    for kb in (96, 112, 128, 144):
        setenv(SPL_HOT_KB=kb)
        A = sp.CsrMatrix.from_device_arrays(n, n, nnz, pp, pi, pv, np.float32, validate=False, ctx=ctx)
        for threads in (1024, 768):
            if kb * 1024 + threads * 32 > 227 * 1024:
                continue
            setenv(SPL_HOT_THREADS=threads)
            try:
                ms = rate([A], [x], [y], 3, 0, reps=10, warm=3)       # the second product builds the table
            except Exception as e:                                   # noqa: BLE001
                res[f"hot{kb}_t{threads}"] = {"error": str(e)[:80]}
                continue
            if y_ref is None:
                y_ref = y.clone()
            err = float((y - y_ref).abs().max() / (y_ref.abs().max() + 1e-30))
            res[f"hot{kb}_t{threads}"] = {"ms": round(ms, 4), "frac": round(b / ms / 1e6 / PEAK, 4), "maxrel_vs_first": err}
            print("c4", kb, threads, res[f"hot{kb}_t{threads}"], flush=True)
        del A
        torch.cuda.empty_cache()
    setenv(SPL_HOT_KB=None, SPL_HOT_THREADS=None)
    OUT["c4"] = res


if __name__ == "__main__":
    for w in sys.argv[1:] or ["c1", "c2", "c5", "c3", "c4"]:
        try:
            globals()[w]()
        except Exception as e:                                       # noqa: BLE001
            OUT[w] = {"error": repr(e)[:300]}
            print(w, "FAILED", repr(e)[:300], flush=True)
        torch.cuda.empty_cache()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "r2_spmv_sweep.json"), "w") as f:
            json.dump(OUT, f, indent=1)
