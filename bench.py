#!/usr/bin/env python
"""bench.py — the hot path of spalinalg on B200, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step is one SpMV y = A x over the workload.  N = 1: config 2 of BASELINE.json (27-point stencil
128^3, f64, 2.1 M rows, 55.7 M nnz; footprint 702 MB > L2, so no flush is needed).  N > 1: config
5 (banded 9, n = 1e8, f64), row-sharded over the ranks — fixed total work, strong scaling.  The
exchange of x is part of every step: `peer` (default for banded/stencil shards) leaves x in its
owners' peer-visible memory and gathers it inside the SpMV kernel over NVLink, ordered by a
device-side flag barrier; `halo` sends boundary slices with NCCL send/recv; `allgather` is the
NCCL all-gather general matrices need.  `value` = algorithmic bytes (nnz*(4+V) + (ncols+nrows)*V,
SURVEY.md 8d) of the whole job per second of the slowest rank, device-timed.  `e2e` = the same
metric through the reference-facing call `&A * &x` with HOST vectors: A is a constructed
CsrMatrix (device resident, as the reference's is host resident), every step copies x from pinned
host memory to the device, runs the SpMV and copies y back (spl_spmv_host at N = 1; upload into
the peer slice + barrier + spl_spmv_peer + download at N > 1).  `e2e_cold` adds the construction
of A from the reference-layout host arrays (usize indices, validating CsrMatrix::new) to every
step.  The line also carries the secondary metrics of the path (COO->CSR assembly on config 1,
CSR->CSC on config 2, config 5 on one GPU as the strong-scaling base; sharded add and sharded
assembly at N > 1), `roofline`, `cpu_baseline` (oracle port of the reference's only SpMV route,
`&A * &X` with X n x 1, on a bounded sample) and `clocks`.

--impl reference times that CPU route alone (the reference is Rust; no rustc here, so the
oracle's line-by-line port stands in: cpu_baseline.kind = "port").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

I_BYTES = 4   # device index width used for the algorithmic byte count


def spmv_bytes(nnz, nrows, ncols, v):
    return nnz * (I_BYTES + v) + (ncols + nrows) * v


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- CPU side
def cpu_sample(kind="stencil", m=64):
    """Bounded sample of the bench workload for the CPU legs: the same stencil at m^3."""
    from spalinalg_b200 import synthetic as syn
    r, c, v = syn.stencil_27(m)
    n = m ** 3
    ptr, ind, val = syn.csr_from_sorted_triplets(n, r, c, v)
    x = 1.0 / (1.0 + (np.arange(n) % 97))
    return n, ptr, ind, val, x


def cpu_reference_step(n, ptr, ind, val, x):
    """The reference's only SpMV route: &A * &X with X n x 1 (src/csr/ops/mul.rs:5-60)."""
    import oracle as orc
    xs = (np.arange(n + 1, dtype=np.uint64), np.zeros(n, np.uint64), x)
    t0 = time.perf_counter()
    orc.csr_mul(n, n, 1, (ptr, ind, val), xs, cap=n)      # n x 1 result: at most n entries
    return time.perf_counter() - t0


def cpu_assembly_baseline():
    """COO->CSR on the CPU: the oracle port of From<&CooMatrix> (src/csr/conv/coo.rs:3-116) on a 512^2
    Laplacian with the triplets shuffled like the device run's.  One core: the reference is
    single-threaded."""
    import oracle as orc
    from spalinalg_b200 import synthetic as syn
    ar, ac, av = syn.laplacian_2d(512)
    perm = np.random.default_rng(42).permutation(len(av))
    trip = orc.make_triplets(ar[perm], ac[perm], av[perm])
    orc.compress_from_coo(512 * 512, 512 * 512, trip, "row")
    t0 = time.perf_counter()
    for _ in range(3):
        orc.compress_from_coo(512 * 512, 512 * 512, trip, "row")
    asm_s = (time.perf_counter() - t0) / 3
    return {"value": len(av) / asm_s / 1e6, "unit": "Mnnz/s", "cores": 1,
            "sample": f"2-D Laplacian 512^2, {len(av)} shuffled triplets, f64, COO->CSR"}


def cpu_mul_baseline():
    """A * A on the CPU: the oracle port of impl Mul for &CsrMatrix (src/csr/ops/mul.rs:5-60: three
    transposes + Gustavson) on a 256^2 Laplacian, one core."""
    import oracle as orc
    from spalinalg_b200 import synthetic as syn
    r, c, v = syn.laplacian_2d(256)
    nn = 256 * 256
    a = orc.compress_from_coo(nn, nn, orc.make_triplets(r, c, v), "row")
    out = orc.csr_mul(nn, nn, nn, a, a)
    t0 = time.perf_counter()
    for _ in range(3):
        out = orc.csr_mul(nn, nn, nn, a, a)
    s_ = (time.perf_counter() - t0) / 3
    return {"value": len(out[1]) / s_ / 1e6, "unit": "Mnnz_out/s", "cores": 1,
            "sample": f"A * A, 2-D Laplacian 256^2, nnz(C)={len(out[1])}, f64"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = 64
    n, ptr, ind, val, x = cpu_sample(m=m)
    nnz = len(val)
    b = spmv_bytes(nnz, n, n, 8)
    for _ in range(args.warmup):
        cpu_reference_step(n, ptr, ind, val, x)
    t = [cpu_reference_step(n, ptr, ind, val, x) for _ in range(args.steps)]
    per = sum(t) / len(t)
    value = b / per / 1e9
    sample = f"27-point stencil {m}^3 (n={n}, nnz={nnz}), f64, X n x 1"
    asm = cpu_assembly_baseline()
    line = {
        "impl": "reference", "metric": "spmv_algorithmic_bandwidth", "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(), "assembly": asm, "mul": cpu_mul_baseline()},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    if args.workload != "auto":
        return args.workload
    return "stencil27_128_f64" if args.gpus == 1 else "banded9_1e8_f64"


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for k, nme in enumerate(names):
                    if r[4 + k].strip().lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        # "under load": drop samples taken while the GPU was idling (power well below the peak seen)
        if power:
            thr = 0.6 * max(power)
            load = [s for s, p in zip(sm, power) if p >= thr] or sm
        else:
            load = sm
        return {"sm_mhz": statistics.median(load) if load else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(load),
                "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto",
                    choices=["auto", "stencil27_128_f64", "laplace2d_1024_f64", "banded9_1e8_f64", "banded9_1e7_f64"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "halo", "allgather"])
    ap.add_argument("--kernel", default="auto")       # auto | vectorN | merge
    ap.add_argument("--no-extras", action="store_true", help="skip secondary metrics / cpu baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import spalinalg_b200 as sp
    from spalinalg_b200 import _capi as capi, sharding
    from spalinalg_b200 import synthetic_device
    from spalinalg_b200.synthetic_device import banded_device, stencil_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    ngpu = world

    # one explicit stream for everything: our kernels, torch's generators, NCCL and the timing
    # events (torch's default stream has handle 0, which the C ABI reads as "create your own")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = sp.Context(local, stream.cuda_stream)
    sp.set_default_context(ctx)
    lib = ctx._lib
    wl = workload_name(args)
    f64 = torch.float64
    V = 8

    # ---- build the workload on the device -------------------------------------------------
    halo = 0
    if wl.startswith("stencil27"):
        offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
        n, rowptr, colind, values = stencil_device(torch, offs, 128, 26.0, -1.0, f64)
        r0, r1 = 0, n
    elif wl.startswith("laplace2d"):
        offs = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]
        n, rowptr, colind, values = stencil_device(torch, offs, 1024, 4.0, -1.0, f64)
        r0, r1 = 0, n
    else:
        n = 10 ** 8 if "1e8" in wl else 10 ** 7
        r0, r1 = sharding.row_partition(n, ngpu, rank)
        rowptr, colind, values = banded_device(torch, n, r0, r1, range(-4, 5), f64)
        halo = 4
    nloc = r1 - r0
    nnz_loc = int(colind.numel())
    A = sp.CsrMatrix.from_device_arrays(nloc, n, nnz_loc, rowptr.data_ptr(), colind.data_ptr(),
                                        values.data_ptr(), np.float64, validate=True, ctx=ctx)
    nnz_t = torch.tensor([nnz_loc], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())
    bytes_total = spmv_bytes(nnz_total, n, n, V)
    bytes_local = spmv_bytes(nnz_loc, nloc, n if ngpu == 1 else nloc + 2 * halo, V)

    # x: `peer` keeps only the owned slice (peer-visible memory, mapped by every rank); the NCCL
    # exchanges use a global-length buffer per rank of which the rank owns [r0, r1)
    from spalinalg_b200 import dist as spd
    exchange = "none"
    if world > 1:
        exchange = args.exchange if args.exchange != "auto" else ("peer" if halo else "allgather")

    def x_values(lo, hi):
        idx = torch.arange(lo, hi, device="cuda", dtype=torch.int64)
        if wl.startswith("banded"):
            return torch.sin(idx.to(f64) * 1e-3)
        return (1.0 / (1.0 + (idx % 97))).to(f64)

    xv = None
    if exchange == "peer":
        starts = spd.partition_starts(n, world)
        xv = spd.PeerVector(ctx, dist, n, np.float64, starts)
        x_loc = synthetic_device.device_view(torch, xv.local_ptr, nloc, f64)
        x_loc.copy_(x_values(r0, r1))
        x_full = None
        dA = spd.DistCsrMatrix(A, starts, rank, n, n)
    else:
        x_full = x_values(0, n)
        x_loc = x_full[r0:r1]
    y = torch.zeros(nloc, device="cuda", dtype=f64)

    kern, lanes = capi.SPL_SPMV_AUTO, 0
    if args.kernel.startswith("vector"):
        kern, lanes = capi.SPL_SPMV_VECTOR, int(args.kernel[6:] or 0)
    elif args.kernel == "merge":
        kern = capi.SPL_SPMV_MERGE

    def local_spmv():
        if exchange == "peer":
            dA.spmv_peer(xv, y.data_ptr())
        else:
            A.spmv_device(x_full.data_ptr(), y.data_ptr(), kern, lanes)

    def step():
        if exchange == "peer":
            xv.barrier()                      # device-side: peers' slices are final before the gathers
        elif exchange == "allgather":
            sharding.exchange_allgather(dist, x_full, r0, r1, world, n % world == 0)
        elif exchange == "halo":
            sharding.exchange_halo(dist, x_full, r0, r1, halo, rank, world)
        local_spmv()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness spot check before timing: analytic row sums are checked by the tests; here
    # the sharded result must equal the single-buffer kernel's on the same rows ----------------
    step()
    torch.cuda.synchronize()
    if exchange == "peer":
        xv.check()
        lo, hi = max(0, r0 - halo), min(n, r1 + halo)
        x_chk = torch.zeros(n if n <= 2 * 10 ** 8 else 1, device="cuda", dtype=f64)
        x_chk[lo:hi] = x_values(lo, hi)
        y_chk = torch.empty_like(y)
        A.spmv_device(x_chk.data_ptr(), y_chk.data_ptr(), kern, lanes)
        torch.cuda.synchronize()
        assert torch.equal(y, y_chk), "peer-gather SpMV differs from the single-buffer kernel"
        del x_chk, y_chk

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    # ---- device-timed steps ----------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = bytes_total / (ms_step * 1e-3) / 1e9
    if xv is not None:
        xv.check()

    # SpMV-kernel-only time on this rank (roofline of the dominant kernel)
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        local_spmv()
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / args.steps
    peak, peak_src = peaks()
    achieved = bytes_local / (kern_ms * 1e-3) / 1e9
    choice = A.spmv_choice()
    kname = "vector" if exchange == "peer" else {1: "vector", 2: "merge"}.get(kern or choice[0], "?")
    traffic = None
    try:                                   # per-launch DRAM bytes from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t_ = json.load(f)[wl]["spmv_" + kname]
            traffic = t_["dram_read_bytes"] + t_["dram_write_bytes"]
    except Exception:
        pass

    # ---- e2e through the reference-facing call with HOST vectors ------------------------------
    # A is a constructed CsrMatrix (device resident; the reference's lives in host memory); a step
    # is `&A * &x`: x from pinned host memory, SpMV, y back to pinned host memory.
    e2e_steps = max(3, min(args.steps, 20))
    h_x = x_loc.cpu().pin_memory() if world > 1 else x_full.cpu().pin_memory()
    h_y = torch.empty(nloc, dtype=f64).pin_memory()
    h2d = h_x.numel() * 8
    d2h = h_y.numel() * 8

    def e2e_step():
        if world == 1:
            ctx.check(lib.spl_spmv_host(ctx._h, A._h, C.c_void_p(h_x.data_ptr()), C.c_void_p(h_y.data_ptr())))
            return
        x_loc.copy_(h_x, non_blocking=True)           # this rank's slice of the new x
        step()                                        # exchange + SpMV
        h_y.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    assert torch.equal(h_y, y.cpu()), "e2e result differs from the device-resident result"
    barrier()
    e2e_l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_launches = (ctx.launch_count() - e2e_l0) // e2e_steps
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = bytes_total / float(t.item()) / 1e9

    # cold variant: construction of A from the reference-layout host arrays inside every step
    cold = None
    if world == 1 and not args.no_extras:
        h_ptr = rowptr.to(torch.int64).cpu().pin_memory()
        h_ind = colind.to(torch.int64).cpu().pin_memory()
        h_val = values.cpu().pin_memory()

        def cold_step():
            h = C.c_void_p()
            ctx.check(lib.spl_mat_from_compressed(ctx._h, capi.SPL_CSR, capi.SPL_F64, nloc, n,
                                                  h_ptr.numel(), C.c_void_p(h_ptr.data_ptr()),
                                                  h_ind.numel(), C.c_void_p(h_ind.data_ptr()),
                                                  h_val.numel(), C.c_void_p(h_val.data_ptr()), C.byref(h)))
            ctx.check(lib.spl_spmv_host(ctx._h, h, C.c_void_p(h_x.data_ptr()), C.c_void_p(h_y.data_ptr())))
            ctx.check(lib.spl_mat_free(ctx._h, h))

        cold_step()
        t0 = time.perf_counter()
        for _ in range(3):
            cold_step()
        torch.cuda.synchronize()
        cs = (time.perf_counter() - t0) / 3
        cold = {"value": bytes_total / cs / 1e9, "unit": "GB/s", "ms_per_step": cs * 1e3,
                "h2d_bytes_per_step": (h_ptr.numel() + h_ind.numel() + h_val.numel()) * 8 + h2d,
                "what": "CsrMatrix::new from host usize arrays (validating) + &A * &x, every step"}
        del h_ptr, h_ind, h_val

    # keep the device busy long enough for a few clock samples if the timed region was short
    if rank == 0:
        t_end = time.perf_counter() + max(0.0, 0.6 - ms_total * 2e-3)
        while time.perf_counter() < t_end:
            for _ in range(50):
                local_spmv()
            torch.cuda.synchronize()
        clk = clocks.stop()

    # ---- secondary metrics of the path ------------------------------------------------------
    extras = {}
    cpu = None
    if ngpu > 1 and not args.no_extras:
        extras = sharded_extras(torch, dist, sp, spd, ctx, A, n, r0, r1, rank, world)
    if rank == 0 and ngpu == 1 and not args.no_extras:
        extras = secondary_metrics(torch, sp, ctx, A, wl)
        m = 64
        sn, sptr, sind, sval, sx = cpu_sample(m=m)
        sb = spmv_bytes(len(sval), sn, sn, 8)
        cpu_reference_step(sn, sptr, sind, sval, sx)
        ts = [cpu_reference_step(sn, sptr, sind, sval, sx) for _ in range(3)]
        import oracle as orc
        t0 = time.perf_counter()
        for _ in range(5):
            orc.csr_spmv(sn, sptr, sind, sval, sx)
        rowdot = (time.perf_counter() - t0) / 5
        cpu = {"value": sb / statistics.median(ts) / 1e9, "unit": "GB/s", "cores": 1,
               "assembly": cpu_assembly_baseline(), "mul": cpu_mul_baseline(),
               "kind": "port", "host_cores": os.cpu_count(),
               "sample": f"27-point stencil {m}^3 (n={sn}, nnz={len(sval)}), f64; reference route "
                         f"&A * &X (3 transposes + Gustavson), 3 reps median",
               "rowdot_value": sb / rowdot / 1e9}

    if rank == 0:
        line = {
            "metric": "spmv_algorithmic_bandwidth", "value": value, "unit": "GB/s", "n_gpus": ngpu,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if ngpu > 1 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "nrows": n, "nnz": nnz_total, "algorithmic_bytes": bytes_total,
                       "l2": "inputs larger than L2 (no flush)" if bytes_local > 2 * 126e6 else
                             "footprint below 2x L2: L2-resident number",
                       "exchange": exchange, "spmv_kernel": kname,
                       "lanes_per_row": choice[1] if lanes == 0 else lanes,
                       "pct_of_8TBps_nominal": 100.0 * achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d * ngpu, "d2h_bytes_per_step": d2h * ngpu,
                    "steps": e2e_steps, "launches_per_step": e2e_launches,
                    "what": "&A * &x on a constructed (device-resident) CsrMatrix with pinned host x, y: "
                            + ("spl_spmv_host" if ngpu == 1 else f"slice upload + {exchange} exchange + SpMV + download, all ranks")},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "spmv", "kernel_ms": kern_ms, "bytes_per_launch": bytes_local},
            "clocks": clk,
        }
        if cold:
            line["e2e_cold"] = cold
        if cpu:
            line["cpu_baseline"] = cpu
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sharded_extras(torch, dist, sp, spd, ctx, A, n, r0, r1, rank, world):
    """Row-sharded add (config 5: A + B, B = offsets {-8,-2,0,2,8}; no exchange) and row-sharded
    COO->CSR assembly (config-3 style random triplets, block-distributed by entry index, routed by
    one all-to-all).  Times are the slowest rank's; rates are whole-job."""
    from spalinalg_b200.synthetic_device import banded_device, random_uniform_coo_device
    out = {}

    def timed_all(fn, reps=3, warm=1):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return statistics.median(ts)

    def total(v):
        t = torch.tensor([v], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    bp, bc, bv = banded_device(torch, n, r0, r1, (-8, -2, 0, 2, 8), torch.float64)
    B = sp.CsrMatrix.from_device_arrays(r1 - r0, n, bc.numel(), bp.data_ptr(), bc.data_ptr(), bv.data_ptr(),
                                        np.float64, validate=False, ctx=ctx)
    del bp, bc, bv
    keep = {}

    def add():
        keep["C"] = A + B
    ms = timed_all(add)
    na, nb, nc = total(A.nnz()), total(B.nnz()), total(keep["C"].nnz())
    b_add = (na + nb + nc) * 12 + 3 * (n + world) * 4
    out["sharded_add"] = {"workload": "banded9 + banded{-8,-2,0,2,8}, n=1e8, f64, rows sharded, no exchange",
                          "ms": ms, "nnz_out": nc, "mnnz_per_s": (na + nb) / ms / 1e3,
                          "gbps_algorithmic": b_add / ms / 1e6}
    del keep["C"], B

    nr = 10_000_000
    per = nr // world                      # each rank emits the triplets of `per` rows' worth of entries
    r_, c_, v_ = random_uniform_coo_device(torch, per, 16, per * 16 // 20, torch.float32, seed=100 + rank)
    # spread the rows over the whole matrix so that ~ (world-1)/world of the entries change rank
    r_ = ((r_.to(torch.int64) * world + rank) % nr).to(torch.int32)
    c_ = ((c_.to(torch.int64) * world + (rank * 7) % world) % nr).to(torch.int32)

    def asm():
        keep["D"] = spd.DistCsrMatrix.from_device_triplets(dist, torch, nr, nr, r_, c_, v_, ctx=ctx)
    ms = timed_all(asm)
    ln = total(int(v_.numel()))
    ex = spd.PeerExchange(ctx, dist)

    def asm_peer():
        keep["P"] = spd.DistCsrMatrix.from_device_triplets_peer(dist, torch, nr, nr, r_, c_, v_, ex)
    ms_p = timed_all(asm_peer)
    ex.check()
    assert keep["P"].local.nnz() == keep["D"].local.nnz()
    del keep["P"]
    ex.close()
    out["sharded_assembly_peer"] = {"workload": "same triplets; routing and exchange fused: the partition pass writes "
                                                "into the owners' buffers over NVLink (peer memory), no all-to-all",
                                    "len": total(int(v_.numel())), "ms": ms_p,
                                    "mnnz_per_s": total(int(v_.numel())) / ms_p / 1e3}
    D = keep["D"]
    nnz_d = total(D.local.nnz())
    out["sharded_assembly"] = {"workload": "random 1e7 x 1e7, 16/row + 5% duplicates, f32, triplets block-"
                                           "distributed by entry index, one all-to-all",
                               "len": ln, "nnz_out": nnz_d, "ms": ms, "mnnz_per_s": ln / ms / 1e3}
    del r_, c_, v_

    # general (random-column) matrix: x exchanged by NCCL all-gather, then the local SpMV
    d0, d1 = D.local_rows()
    xg = torch.rand(nr, device="cuda", dtype=torch.float32) - 0.5
    yg = torch.empty(d1 - d0, device="cuda", dtype=torch.float32)
    even = nr % world == 0

    def spmv_ag():
        from spalinalg_b200 import sharding
        sharding.exchange_allgather(dist, xg, d0, d1, world, even)
        D.local.spmv_device(xg.data_ptr(), yg.data_ptr())
    ms = timed_all(spmv_ag, reps=7, warm=3)
    b_ag = nnz_d * 8 + 2 * nr * 4
    # the same exchange done by pulling the peers' slices over NVLink (peer memory) instead of NCCL
    from spalinalg_b200.synthetic_device import device_view
    xs = spd.PeerVector(ctx, dist, nr, np.float32, D.starts)
    device_view(torch, xs.local_ptr, d1 - d0, torch.float32).copy_(xg[d0:d1])
    torch.cuda.synchronize()

    def spmv_pull():
        xs.barrier()
        xs.pull(xg.data_ptr())
        D.local.spmv_device(xg.data_ptr(), yg.data_ptr())
    y_ag = yg.clone()
    ms_pull = timed_all(spmv_pull, reps=7, warm=3)
    xs.check()
    assert torch.equal(y_ag, yg), "pulled all-gather gives a different y"
    out["sharded_spmv_peer_pull"] = {"workload": "same matrix; x all-gathered by one pull kernel over peer memory "
                                                 "(device barrier + 128-bit NVLink reads), then the local SpMV",
                                     "ms": ms_pull, "gbps_algorithmic": b_ag / ms_pull / 1e6}
    xs.close(dist)
    out["sharded_spmv_allgather"] = {"workload": "the matrix assembled above (random 16/row, f32), x all-gathered "
                                                 "with NCCL every step", "ms": ms, "gbps_algorithmic": b_ag / ms / 1e6,
                                     "x_bytes_received_per_rank": (nr - (d1 - d0)) * 4}
    return out if rank == 0 else {}


def secondary_metrics(torch, sp, ctx, A, wl):
    """COO->CSR assembly on config 1 (shuffled and row-ordered), CSR->CSC on the bench matrix, and
    SpMV on config 5 (one GPU: the base of the N > 1 strong-scaling runs) and config 1."""
    from spalinalg_b200.synthetic_device import banded_device, stencil_device
    out = {}

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    # CSR -> CSC / transpose of the bench matrix
    nnz = A.nnz()
    ms = timed(lambda: A.to_csc())
    b_tr = 2 * nnz * (4 + 8) + (A.nrows() + 1) * 4 + (A.ncols() + 1) * 4
    out["csr_to_csc"] = {"workload": wl, "ms": ms, "mnnz_per_s": nnz / ms / 1e3, "gbps_algorithmic": b_tr / ms / 1e6}

    # assembly: config 1, shuffled COO, device-resident SoA in -> device CSR out
    offs = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]
    n, rowptr, colind, values = stencil_device(torch, offs, 1024, 4.0, -1.0, torch.float64)
    rows = torch.repeat_interleave(torch.arange(n, device="cuda", dtype=torch.int32),
                                   (rowptr[1:] - rowptr[:-1]).to(torch.int64))
    g = torch.Generator(device="cuda"); g.manual_seed(42)
    perm = torch.randperm(rows.numel(), device="cuda", generator=g)
    out_host_triplets = (rows[perm].contiguous(), colind[perm].contiguous(), values[perm].contiguous())
    for name, (r_, c_, v_) in {"shuffled": out_host_triplets,
                               "row_ordered": (rows, colind, values)}.items():
        ln = r_.numel()
        ms = timed(lambda: sp.CsrMatrix.from_device_triplets(n, n, ln, r_.data_ptr(), c_.data_ptr(),
                                                             v_.data_ptr(), np.float64, ctx=ctx))
        b_asm = ln * (8 + 8) + ln * (4 + 8) + (n + 1) * 4
        out[f"assembly_{name}"] = {"workload": "laplace2d_1024_f64 COO->CSR", "len": ln, "ms": ms,
                                   "mnnz_per_s": ln / ms / 1e3, "gbps_algorithmic": b_asm / ms / 1e6}

    # the same assembly through the reference-facing call: host usize (u64) row/col + f64 values in
    # pinned memory -> spl_mat_from_coo (upload, narrow, sort, sum) -> CsrMatrix
    sr, sc_, sv = out_host_triplets
    coo_h = [t.cpu().pin_memory() for t in (sr.to(torch.int64), sc_.to(torch.int64), sv)]
    lib = ctx._lib
    import ctypes as C
    from spalinalg_b200 import _capi as capi

    def asm_host():
        h = C.c_void_p()
        ctx.check(lib.spl_mat_from_coo(ctx._h, capi.SPL_CSR, capi.SPL_F64, n, n, coo_h[2].numel(),
                                       C.c_void_p(coo_h[0].data_ptr()), C.c_void_p(coo_h[1].data_ptr()),
                                       C.c_void_p(coo_h[2].data_ptr()), 1, 1, C.byref(h)))
        ctx.check(lib.spl_mat_free(ctx._h, h))
    asm_host(); asm_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        asm_host()
    torch.cuda.synchronize()
    hs = (time.perf_counter() - t0) / 5
    out["assembly_shuffled_e2e"] = {"workload": "laplace2d_1024_f64 COO->CSR from pinned host usize triplets",
                                    "len": int(coo_h[2].numel()), "ms": hs * 1e3,
                                    "mnnz_per_s": coo_h[2].numel() / hs / 1e6,
                                    "h2d_bytes": int(coo_h[2].numel()) * 24}

    # the same triplets through the streaming CooMatrix storage (spl_coo): the fill
    # (extend from the host arrays, chunks sent while the rest is copied in) and the conversion
    # (flush of the last partial chunk + device assembly) timed separately and together
    np_trip = [t.numpy() for t in coo_h]
    np_trip = (np_trip[0].view(np.uint64), np_trip[1].view(np.uint64), np_trip[2])
    ln = int(np_trip[2].shape[0])
    fill_s, conv_s = [], []
    for it in range(6):
        pc = sp.PinnedCooMatrix.with_capacity(n, n, ln, np.float64, ctx=ctx)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pc.extend_triplets(*np_trip)
        t1 = time.perf_counter()
        M_ = sp.CsrMatrix.from_coo(pc, ctx=ctx)
        ctx.sync()
        t2 = time.perf_counter()
        if it:
            fill_s.append(t1 - t0); conv_s.append(t2 - t1)
        del M_, pc
    fill_ms, conv_ms = 1e3 * float(np.median(fill_s)), 1e3 * float(np.median(conv_s))
    out["assembly_shuffled_streamed"] = {
        "workload": "laplace2d_1024_f64: PinnedCooMatrix filled from host usize triplets, then CsrMatrix::from(&coo)",
        "len": ln, "fill_ms": fill_ms, "convert_ms": conv_ms, "total_ms": fill_ms + conv_ms,
        "convert_mnnz_per_s": ln / conv_ms / 1e3, "total_mnnz_per_s": ln / (fill_ms + conv_ms) / 1e3}
    del coo_h, np_trip

    def spmv_rate(M, x, y, copies=1, reps=100):
        """Back-to-back launches; `copies` > 1 rotates over independent (A, x, y) sets whose total
        footprint exceeds L2, so every launch reads from HBM."""
        for i in range(5):
            M[i % copies].spmv_device(x[i % copies].data_ptr(), y[i % copies].data_ptr())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(reps):
            M[i % copies].spmv_device(x[i % copies].data_ptr(), y[i % copies].data_ptr())
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # the reference's Mul on a general right-hand side: A * A on the config-1 matrix
    A1m = sp.CsrMatrix.from_device_arrays(n, n, colind.numel(), rowptr.data_ptr(), colind.data_ptr(),
                                          values.data_ptr(), np.float64, validate=False, ctx=ctx)
    keepm = {}

    def mul():
        keepm["C"] = A1m * A1m
    ms = timed(mul, reps=5, warm=2)
    out["mul_laplace2d_1024_f64"] = {"workload": "A * A, 2-D Laplacian 1024^2 (impl Mul for &CsrMatrix)", "ms": ms,
                                     "nnz_out": keepm["C"].nnz(), "mnnz_out_per_s": keepm["C"].nnz() / ms / 1e3}
    del A1m, keepm

    peak, _ = peaks()
    # config 1 SpMV, 4 rotating copies (4 x 80 MB > L2)
    A1 = [sp.CsrMatrix.from_device_arrays(n, n, colind.numel(), rowptr.data_ptr(), colind.data_ptr(),
                                          values.data_ptr(), np.float64, validate=False, ctx=ctx) for _ in range(4)]
    x1 = [torch.rand(n, device="cuda", dtype=torch.float64) for _ in range(4)]
    y1 = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(4)]
    b1 = spmv_bytes(int(colind.numel()), n, n, 8)
    ms = spmv_rate(A1, x1, y1, copies=4, reps=200)
    out["spmv_laplace2d_1024_f64"] = {"ms": ms, "gbps": b1 / ms / 1e6, "frac_measured_peak": b1 / ms / 1e6 / peak,
                                      "l2": "4 rotating copies of (A, x, y), 320 MB > L2"}
    ms = spmv_rate(A1[:1], x1[:1], y1[:1], reps=200)
    out["spmv_laplace2d_1024_f64_l2_resident"] = {"ms": ms, "gbps": b1 / ms / 1e6}
    del A1, x1, y1, rows, perm, out_host_triplets
    # config 5 on one GPU
    n5 = 10 ** 8
    p5, c5, v5 = banded_device(torch, n5, 0, n5, range(-4, 5), torch.float64)
    A5 = sp.CsrMatrix.from_device_arrays(n5, n5, c5.numel(), p5.data_ptr(), c5.data_ptr(), v5.data_ptr(),
                                         np.float64, validate=False, ctx=ctx)
    nnz5 = int(c5.numel())
    del p5, c5, v5
    x5 = torch.rand(n5, device="cuda", dtype=torch.float64)
    y5 = torch.empty(n5, device="cuda", dtype=torch.float64)
    b5 = spmv_bytes(nnz5, n5, n5, 8)
    ms = spmv_rate([A5], [x5], [y5], reps=20)
    out["spmv_banded9_1e8_f64_1gpu"] = {"ms": ms, "gbps": b5 / ms / 1e6, "frac_measured_peak": b5 / ms / 1e6 / peak,
                                        "note": "same workload as the N > 1 runs: base of their strong scaling"}
    return out


if __name__ == "__main__":
    main()
