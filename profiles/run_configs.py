#!/usr/bin/env python
"""Measures every BASELINE.json config on one B200 (device-resident, CUDA events, median of reps):
assembly Mnnz/s, CSR->CSC, transpose, SpMV per kernel with algorithmic GB/s and fraction of the
measured HBM peak, add.  Writes gpurun_out/configs.json.  Usage: python profiles/run_configs.py [c1 c2 ...]"""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spalinalg_b200 as sp                                        # noqa: E402
from spalinalg_b200 import _capi as capi                           # noqa: E402
from spalinalg_b200 import synthetic_device as sd                  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sp.Context(0, stream.cuda_stream)
sp.set_default_context(ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > L2 (126 MB)


def timed(fn, reps=7, warm=2, flush_l2=False):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


def spmv_rows(A, V, tdt, small):
    n, m, nnz = A.nrows(), A.ncols(), A.nnz()
    x = torch.rand(m, device="cuda", dtype=tdt) - 0.5
    y = torch.empty(n, device="cuda", dtype=tdt)
    b = nnz * (4 + V) + (n + m) * V
    out = {"algorithmic_bytes": b, "planned": A.spmv_choice()}
    variants = [("auto", 0, 0), ("merge", 2, 0), ("split", 3, 0), ("sliced", 4, 0)] + [(f"vector{l}", 1, l) for l in (1, 2, 4, 8, 16, 32)]
    for name, k, l in variants:
        try:
            med, mn = timed(lambda: A.spmv_device(x.data_ptr(), y.data_ptr(), k, l), reps=15, warm=3, flush_l2=small)
        except Exception as e:                                       # noqa: BLE001
            out[name] = {"error": str(e)[:80]}
            continue
        out[name] = {"ms": med, "gbps": b / med / 1e6, "frac_measured_peak": b / med / 1e6 / PEAK,
                     "frac_8TBps": b / med / 1e6 / 8000.0}
    return out


def tr_bytes(A, V):
    return 2 * A.nnz() * (4 + V) + (A.nrows() + 1) * 4 + (A.ncols() + 1) * 4


def assembly(n, m, r, c, v, npdt, V, fmt="csr"):
    cls = sp.CsrMatrix if fmt == "csr" else sp.CscMatrix
    ln = r.numel()
    keep = {}

    def run():
        keep["A"] = cls.from_device_triplets(n, m, ln, r.data_ptr(), c.data_ptr(), v.data_ptr(), npdt, ctx=ctx)
    med, mn = timed(run, reps=5, warm=1)
    A = keep["A"]
    b = ln * (8 + V) + A.nnz() * (4 + V) + (n + 1) * 4
    return A, {"len": ln, "nnz_out": A.nnz(), "ms": med, "mnnz_per_s": ln / med / 1e3,
               "gbps_algorithmic": b / med / 1e6, "frac_measured_peak": b / med / 1e6 / PEAK}


def convert_rows(A, V):
    out = {}
    med, _ = timed(lambda: A.to_csc(), reps=5, warm=1)
    out["csr_to_csc"] = {"ms": med, "mnnz_per_s": A.nnz() / med / 1e3, "gbps_algorithmic": tr_bytes(A, V) / med / 1e6,
                         "frac_measured_peak": tr_bytes(A, V) / med / 1e6 / PEAK}
    Cc = A.to_csc()
    med, _ = timed(lambda: Cc.to_csr(), reps=5, warm=1)
    out["csc_to_csr"] = {"ms": med, "mnnz_per_s": A.nnz() / med / 1e3}
    med, _ = timed(lambda: A.transpose(), reps=5, warm=1)
    out["transpose"] = {"ms": med, "mnnz_per_s": A.nnz() / med / 1e3}
    return out


def shuffled(r, c, v, seed=42):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    p = torch.randperm(r.numel(), device="cuda", generator=g)
    return r[p].contiguous(), c[p].contiguous(), v[p].contiguous()


def rows_of(ptr, n):
    return torch.repeat_interleave(torch.arange(n, device="cuda", dtype=torch.int32), (ptr[1:] - ptr[:-1]).long())


def c1():
    offs = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]
    n, ptr, col, val = sd.stencil_device(torch, offs, 1024, 4.0, -1.0, torch.float64)
    rows = rows_of(ptr, n)
    res = {"config": "C1 2-D Laplacian 1024^2 f64", "nrows": n}
    A, res["assembly_row_ordered"] = assembly(n, n, rows, col, val, np.float64, 8)
    _, res["assembly_shuffled"] = assembly(n, n, *shuffled(rows, col, val), np.float64, 8)
    # row by row, but the columns of every row in descending order (rows sorted, columns not)
    pos = torch.arange(col.numel(), device="cuda", dtype=torch.int64)
    first, last = ptr[:-1].long()[rows.long()], ptr[1:].long()[rows.long()] - 1
    rev = first + (last - pos)
    _, res["assembly_rows_sorted_cols_reversed"] = assembly(n, n, rows, col[rev].contiguous(), val[rev].contiguous(),
                                                            np.float64, 8)
    _, res["assembly_shuffled_csc"] = assembly(n, n, *shuffled(rows, col, val), np.float64, 8, "csc")
    res["spmv_l2_flushed"] = spmv_rows(A, 8, torch.float64, small=True)
    res["spmv_l2_resident"] = spmv_rows(A, 8, torch.float64, small=False)
    res.update(convert_rows(A, 8))
    return res


def c2():
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
    n, ptr, col, val = sd.stencil_device(torch, offs, 128, 26.0, -1.0, torch.float64)
    A = sp.CsrMatrix.from_device_arrays(n, n, col.numel(), ptr.data_ptr(), col.data_ptr(), val.data_ptr(), np.float64, ctx=ctx)
    res = {"config": "C2 27-pt stencil 128^3 f64", "nrows": n, "nnz": A.nnz()}
    res["spmv"] = spmv_rows(A, 8, torch.float64, small=False)
    res.update(convert_rows(A, 8))
    rows = rows_of(ptr, n)
    _, res["assembly_shuffled"] = assembly(n, n, *shuffled(rows, col, val), np.float64, 8)
    return res


def c3():
    n = 10_000_000
    r, c, v = sd.random_uniform_coo_device(torch, n, 16, 8_000_000, torch.float32, seed=1)
    res = {"config": "C3 random 1e7 x 1e7, 16/row + 5% duplicates, f32", "nrows": n}
    A, res["assembly"] = assembly(n, n, r, c, v, np.float32, 4)
    del r, c, v
    res["nnz"] = A.nnz()
    res["spmv"] = spmv_rows(A, 4, torch.float32, small=False)
    res.update(convert_rows(A, 4))
    return res


def c4():
    scale = 24
    n = 1 << scale
    r, c, v = sd.rmat_coo_device(torch, scale, 32, torch.float32, seed=3)
    res = {"config": "C4 R-MAT 2^24, 32 edges/row, f32", "nrows": n}
    A, res["assembly"] = assembly(n, n, r, c, v, np.float32, 4)
    del r, c, v
    res["nnz"] = A.nnz()
    res["spmv"] = spmv_rows(A, 4, torch.float32, small=False)
    res.update(convert_rows(A, 4))
    return res


def c5():
    n = 100_000_000
    ptr, col, val = sd.banded_device(torch, n, 0, n, range(-4, 5), torch.float64)
    A = sp.CsrMatrix.from_device_arrays(n, n, col.numel(), ptr.data_ptr(), col.data_ptr(), val.data_ptr(), np.float64, ctx=ctx)
    del ptr, col, val
    res = {"config": "C5 banded 9, n = 1e8, f64 (one GPU)", "nrows": n, "nnz": A.nnz()}
    res["spmv"] = spmv_rows(A, 8, torch.float64, small=False)
    bp, bc, bv = sd.banded_device(torch, n, 0, n, (-8, -2, 0, 2, 8), torch.float64)
    B = sp.CsrMatrix.from_device_arrays(n, n, bc.numel(), bp.data_ptr(), bc.data_ptr(), bv.data_ptr(), np.float64, ctx=ctx)
    del bp, bc, bv
    keep = {}

    def add():
        keep["C"] = A + B
    med, _ = timed(add, reps=3, warm=1)
    C = keep["C"]
    b = (A.nnz() + B.nnz() + C.nnz()) * 12 + 3 * (n + 1) * 4
    res["add"] = {"ms": med, "nnz_c": C.nnz(), "mnnz_per_s": (A.nnz() + B.nnz()) / med / 1e3,
                  "gbps_algorithmic": b / med / 1e6, "frac_measured_peak": b / med / 1e6 / PEAK}
    return res


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    out = {}
    for w in which:
        out[w] = globals()[w]()
        torch.cuda.empty_cache()
        print(w, json.dumps(out[w])[:3000], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)
