"""The row-sharded path on real GPUs: one process per GPU over NCCL (skipped on a one-GPU box; run
with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).  Every rank checks its
shard against the CPU oracle's result on the WHOLE problem:
  * sharded COO->CSR assembly (device routing, NCCL all-to-all, packed assembly): bit-exact shard;
  * peer-memory SpMV (CUDA IPC slices, device-side barrier, gather in the kernel) and the NCCL
    all-gather variant: within 1e-12 relative (f64);
  * sharded add / sub / neg on the shared partition: bit-exact;
  * row-sharded CSR -> column-sharded CSC (all-to-all by column owner): bit-exact."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import oracle as orc
    import spalinalg_b200 as sp
    from spalinalg_b200 import dist as spd, sharding
    from tests.test_dist import make_coo, shard_of, syn_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    try:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        ctx = sp.Context(rank, stream.cuda_stream)
        sp.set_default_context(ctx)
        n = 20011                                    # not divisible by the world size
        r, c, v = make_coo(n, n, 400000, 123)
        full = orc.compress_from_coo(n, n, orc.make_triplets(r, c, v), "row")
        a, b = syn_block(len(v), world, rank)
        dev = lambda arr, dt: torch.from_numpy(np.ascontiguousarray(arr).astype(dt)).cuda()
        D = spd.DistCsrMatrix.from_device_triplets(dist, torch, n, n, dev(r[a:b], np.int32), dev(c[a:b], np.int32),
                                                   dev(v[a:b], np.float64), ctx=ctx)
        want = shard_of(full, D.starts, rank)
        L = D.local
        ok &= bool(np.array_equal(L.rowptr(), want[0]) and np.array_equal(L.colind(), want[1])
                   and L.values().tobytes() == want[2].tobytes())
        assert ok, "sharded assembly differs from the oracle"
        # the same with routing and exchange fused over peer memory (two calls: buffer reuse, f32 after f64)
        ex = spd.PeerExchange(ctx, dist)
        for dt in (np.float64, np.float32, np.float64):
            vv = v.astype(dt)
            fullp = orc.compress_from_coo(n, n, orc.make_triplets(r, c, vv), "row")
            P = spd.DistCsrMatrix.from_device_triplets_peer(dist, torch, n, n, dev(r[a:b], np.int32),
                                                            dev(c[a:b], np.int32), dev(vv[a:b], dt), ex)
            wp = shard_of(fullp, P.starts, rank)
            ok &= bool(np.array_equal(P.local.rowptr(), wp[0]) and np.array_equal(P.local.colind(), wp[1])
                       and P.local.values().tobytes() == wp[2].tobytes())
        ex.check()
        assert ok, "peer-memory sharded assembly differs from the oracle"
        ex.close()

        # SpMV: x sharded like the columns, left in its owners' memory
        x = np.random.default_rng(9).standard_normal(n)
        yw = orc.csr_spmv(n, *full, x)
        sc = orc.csr_spmv(n, full[0], full[1], np.abs(full[2]), np.abs(x))
        r0, r1 = D.local_rows()
        xv = spd.PeerVector(ctx, dist, n, np.float64, D.starts)
        from spalinalg_b200.synthetic_device import device_view
        device_view(torch, xv.local_ptr, r1 - r0, torch.float64).copy_(torch.from_numpy(x[r0:r1]))
        y = torch.zeros(r1 - r0, dtype=torch.float64, device="cuda")
        xv.publish()                                  # barrier + the written buffer becomes the published one
        for _ in range(3):                            # several epochs of the flag barrier, x unchanged
            D.spmv_peer(xv, y.data_ptr())
            xv.barrier()
        torch.cuda.synchronize()
        xv.check()
        ok &= bool(np.all(np.abs(y.cpu().numpy() - yw[r0:r1]) <= 1e-12 * sc[r0:r1] + 1e-300))
        assert ok, "peer-memory SpMV differs from the oracle"
        x_full = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
        x_full[r0:r1] = torch.from_numpy(x[r0:r1]).cuda()
        sharding.exchange_allgather(dist, x_full, r0, r1, world, n % world == 0)
        y.zero_()
        L.spmv_device(x_full.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        ok &= bool(np.all(np.abs(y.cpu().numpy() - yw[r0:r1]) <= 1e-12 * sc[r0:r1] + 1e-300))
        assert ok, "all-gather SpMV differs from the oracle"
        # the all-gather done by pulling the peers' slices over NVLink (spl_peer_pull)
        x_full.fill_(float("nan"))
        x_full[r0:r1] = torch.from_numpy(x[r0:r1]).cuda()
        torch.cuda.synchronize()
        xv.barrier()
        xv.pull(x_full.data_ptr())                    # pulls the published buffer
        y.zero_()
        L.spmv_device(x_full.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        xv.check()
        ok &= bool(np.all(np.abs(y.cpu().numpy() - yw[r0:r1]) <= 1e-12 * sc[r0:r1] + 1e-300))
        assert ok, "pulled all-gather SpMV differs from the oracle"

        # general shards: the all-gather fused into the product (a copy warp per CTA pulls the slices over NVLink
        # while the consumer warps work through the owner blocks); two products on one set of flags, x changed between
        D.prepare_gather(torch)
        for scale_x in (1.0, -0.5):
            device_view(torch, xv.local_ptr, r1 - r0, torch.float64).copy_(torch.from_numpy(scale_x * x[r0:r1]))
            xv.publish()
            x_full.fill_(float("nan"))
            y.fill_(7.0)
            torch.cuda.synchronize()
            D.spmv_gather(xv, x_full.data_ptr(), y.data_ptr())
            torch.cuda.synchronize()
            ok &= bool(np.all(np.abs(y.cpu().numpy() - scale_x * yw[r0:r1]) <= 4e-12 * sc[r0:r1] + 1e-300))
        xv.check()
        assert ok, "fused gather SpMV differs from the oracle"
        # the same with the barrier folded into the kernel: write the next slice, swap, one launch
        for scale_x in (2.0, -1.0):
            device_view(torch, xv.local_ptr, r1 - r0, torch.float64).copy_(torch.from_numpy(scale_x * x[r0:r1]))
            xv.swap()
            x_full.fill_(float("nan"))
            y.fill_(7.0)
            torch.cuda.synchronize()
            D.spmv_gather(xv, x_full.data_ptr(), y.data_ptr(), barrier=True)
            torch.cuda.synchronize()
            ok &= bool(np.all(np.abs(y.cpu().numpy() - scale_x * yw[r0:r1]) <= 4e-12 * sc[r0:r1] + 1e-300))
        xv.check()
        assert ok, "fused gather SpMV with its own barrier differs from the oracle"
        # f32: rank 1's slice starts at an odd column, so its copy into rank 0's x_full is not 16-byte aligned (the copy
        # warp's element path), and rank 0's slice ends in a tail shorter than 16 bytes behind the bulk copies
        v32 = v.astype(np.float32)
        full32 = orc.compress_from_coo(n, n, orc.make_triplets(r, c, v32), "row")
        D32 = spd.DistCsrMatrix.from_device_triplets(dist, torch, n, n, dev(r[a:b], np.int32), dev(c[a:b], np.int32),
                                                     dev(v32[a:b], np.float32), ctx=ctx)
        x32 = x.astype(np.float32)
        yw32 = orc.csr_spmv(n, *full32, x32).astype(np.float64)
        sc32 = orc.csr_spmv(n, full32[0], full32[1], np.abs(full32[2]), np.abs(x32)).astype(np.float64)
        xv32 = spd.PeerVector(ctx, dist, n, np.float32, D32.starts)
        xf32 = torch.empty(n, dtype=torch.float32, device="cuda")
        y32 = torch.empty(r1 - r0, dtype=torch.float32, device="cuda")
        D32.prepare_gather(torch)
        for scale_x, lanes in ((1.0, None), (-3.0, None), (0.5, "1"), (2.0, "2"), (-1.0, "4")):
            if lanes:                                           # every tile shape of the kernel, not only the one it picks
                os.environ["SPL_GATHER_LANES"] = lanes
                D32.prepare_gather(torch)                       # fresh counters: the shape decides the grid
            device_view(torch, xv32.local_ptr, r1 - r0, torch.float32).copy_(torch.from_numpy(np.float32(scale_x) * x32[r0:r1]))
            xv32.swap()
            xf32.fill_(float("nan"))
            y32.fill_(7.0)
            torch.cuda.synchronize()
            D32.spmv_gather(xv32, xf32.data_ptr(), y32.data_ptr(), barrier=True)
            torch.cuda.synchronize()
            got_x = xf32.cpu().numpy()
            peers = np.ones(n, dtype=bool)
            peers[r0:r1] = False
            ok &= bool(np.array_equal(got_x[peers], (np.float32(scale_x) * x32)[peers]))      # the slices, bit for bit
            ok &= bool(np.all(np.abs(y32.cpu().numpy().astype(np.float64) - scale_x * yw32[r0:r1])
                              <= 1e-5 * abs(scale_x) * sc32[r0:r1] + 1e-30))
        os.environ.pop("SPL_GATHER_LANES", None)
        xv32.check()
        xv32.close(dist)
        assert ok, "fused gather SpMV (f32, unaligned slice) differs from the oracle"

        # an iteration whose x changes every step: y_t is written straight into the unpublished buffer
        # and published as x_{t+1} (one barrier per step).  One rank is held back by a spin kernel at a
        # different point of every step, so the fast rank runs ahead as far as the barrier lets it; with
        # a single buffer it would overwrite a slice its peer is still gathering (write after read).
        # Banded matrix, one lane per row: the sharded iteration must equal the oracle's sequential one
        # bit for bit.
        ok &= iterate_against_oracle(orc, spd, sp, ctx, rank, world, iters=50)
        assert ok, "changing-x iteration over peer memory differs from the oracle's sequential iteration"
        ok &= iterate_against_oracle(orc, spd, sp, ctx, rank, world, iters=50, halo=True)
        assert ok, "changing-x iteration with the halo copied by the barrier kernel differs from the oracle"

        # add / sub / neg on the shared partition
        r2, c2, v2 = make_coo(n, n, 300000, 321)
        full2 = orc.compress_from_coo(n, n, orc.make_triplets(r2, c2, v2), "row")
        a2, b2 = syn_block(len(v2), world, rank)
        E = spd.DistCsrMatrix.from_device_triplets(dist, torch, n, n, dev(r2[a2:b2], np.int32),
                                                   dev(c2[a2:b2], np.int32), dev(v2[a2:b2], np.float64), ctx=ctx)
        for sub, M in ((0, D + E), (1, D - E)):
            w = shard_of(orc.addsub(sub, n, n, full, full2), D.starts, rank)
            ok &= bool(np.array_equal(M.local.rowptr(), w[0]) and np.array_equal(M.local.colind(), w[1])
                       and M.local.values().tobytes() == w[2].tobytes())
        ok &= (-D).local.values().tobytes() == orc.neg(want[2]).tobytes()

        # row-sharded CSR -> column-sharded CSC: the columns of the whole matrix's CSC form
        T = D.to_csc(dist, torch)
        wc = shard_of(orc.recompress(n, n, *full), T.starts, rank)
        ok &= bool(np.array_equal(T.local.colptr(), wc[0]) and np.array_equal(T.local.rowind(), wc[1])
                   and T.local.values().tobytes() == wc[2].tobytes())
        assert ok, "sharded CSR -> CSC differs from the oracle"
        ex2 = spd.PeerExchange(ctx, dist)
        T2 = D.to_csc(dist, torch, exchange=ex2)               # the same through peer memory
        ok &= bool(np.array_equal(T2.local.colptr(), wc[0]) and np.array_equal(T2.local.rowind(), wc[1])
                   and T2.local.values().tobytes() == wc[2].tobytes())
        ex2.check()
        ex2.close()
        assert ok, "peer-memory sharded CSR -> CSC differs from the oracle"
        # y = A x on the column-sharded CSC form (x sharded like the columns; partials summed by reduce-scatter)
        c0, c1 = T.local_cols()
        yt, rst = T.spmv(dist, torch, torch.from_numpy(x[c0:c1]).cuda())
        torch.cuda.synchronize()
        a0, a1 = rst[rank], rst[rank + 1]
        ok &= bool(np.all(np.abs(yt.cpu().numpy() - yw[a0:a1]) <= 4e-12 * sc[a0:a1] + 1e-300))
        assert ok, "column-sharded CSC SpMV differs from the oracle"
        xv.close(dist)
    finally:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
        dist.destroy_process_group()


def iterate_against_oracle(orc, spd, sp, ctx, rank, world, iters=50, n=300_007, delay_cycles=3_000_000, halo=False):
    """x_{t+1} = A x_t over a PeerVector, `iters` steps, against the oracle's sequential iteration.
    Returns True when this rank's slice of the last x equals the oracle's bit for bit."""
    from spalinalg_b200.synthetic_device import device_view
    rows = np.repeat(np.arange(n), 9)
    cols = rows + np.tile(np.arange(-4, 5), n)
    keep = (cols >= 0) & (cols < n)
    rows, cols = rows[keep], cols[keep]
    ptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.uint64)
    val = (0.11 + 0.001 * ((rows * 7 + cols * 3) % 13)).astype(np.float64)     # row sums ~ 1: the iteration stays O(1)
    x0 = np.sin(np.arange(n) * 1e-2) + 1.5
    starts = spd.partition_starts(n, world)
    r0, r1 = starts[rank], starts[rank + 1]
    lo, hi = int(ptr[r0]), int(ptr[r1])
    A = sp.CsrMatrix.new(r1 - r0, n, (ptr[r0:r1 + 1] - ptr[r0]).astype(np.uint64), cols[lo:hi].astype(np.uint64),
                         val[lo:hi], ctx=ctx)
    dA = spd.DistCsrMatrix(A, starts, rank, n, n)
    # halo=True: the halo is copied next to the own slice by the barrier kernel and the product reads one
    # local array (spmv_halo); halo=False: the product gathers from the owners' slices itself (spmv_peer)
    widths = dA.halo_widths(dist, torch) if halo else (0, 0)
    if halo:
        assert widths == (4, 4), widths
    xv = spd.PeerVector(ctx, dist, n, np.float64, starts, halo=widths)
    device_view(torch, xv.local_ptr, r1 - r0, torch.float64).copy_(torch.from_numpy(x0[r0:r1]))
    xv.publish(halo=halo)
    for t in range(iters):
        slow = (t % world) == rank                    # the late rank changes every step
        if slow and t % 2 == 0:
            torch.cuda._sleep(delay_cycles)           # late before its product: peers wait at the barrier
        if halo:
            dA.spmv_halo(xv, xv.local_ptr)            # y_t -> the unpublished buffer
        else:
            dA.spmv_peer(xv, xv.local_ptr)
        if slow and t % 2 == 1:
            torch.cuda._sleep(delay_cycles)           # late after its product, before the barrier
        xv.publish(halo=halo)
    torch.cuda.synchronize()
    xv.check()
    got = device_view(torch, xv.published_ptr, r1 - r0, torch.float64).cpu().numpy()
    ref = x0
    for _ in range(iters):
        ref = orc.csr_spmv(n, ptr, cols.astype(np.uint64), val, ref)
    same = got.tobytes() == ref[r0:r1].tobytes()
    xv.close(dist)
    return bool(same)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("world", [2])
def test_sharded_path_on_real_gpus(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
