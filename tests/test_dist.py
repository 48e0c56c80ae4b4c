"""Row-sharded hot path.

CPU part (not gpu): world-size-2 gloo run of the sharded assembly's host logic — routing counts,
the all-to-all, source-rank-order concatenation — with numpy standing in for the two device calls
and the oracle assembling each shard; the shards must equal the rows of the oracle's assembly of
the whole COO list, bit for bit (duplicates that straddle the two ranks included).

GPU part: the device calls themselves through the C ABI on one GPU, the ranks emulated one after
the other (never as kernels that wait on each other): spl_coo_route_dev against its numpy
statement, spl_mat_from_packed_dev shards against the oracle, spl_spmv_peer with x split over two
buffers against the oracle SpMV, and the flag barrier's arrival / timeout paths."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from spalinalg_b200 import dist as spd
from spalinalg_b200 import synthetic as syn


def make_coo(n, ncols, length, seed, dtype=np.float64):
    """Random triplets with many duplicates (multiplicity >= 3 included) and exact cancellations."""
    rng = np.random.default_rng(seed)
    r = rng.integers(0, n, length).astype(np.uint64)
    c = rng.integers(0, ncols, length).astype(np.uint64)
    v = rng.standard_normal(length).astype(dtype)
    k = length // 4
    src = rng.integers(0, length, k)
    r = np.concatenate([r, r[src], r[src[: k // 2]]])
    c = np.concatenate([c, c[src], c[src[: k // 2]]])
    v = np.concatenate([v, rng.standard_normal(k).astype(dtype), -v[src[: k // 2]]])
    p = rng.permutation(len(v))
    return r[p], c[p], v[p]


def shard_of(full, starts, g):
    ptr, ind, val = full
    a, b = starts[g], starts[g + 1]
    lo, hi = int(ptr[a]), int(ptr[b])
    return (ptr[a:b + 1] - ptr[a]).astype(np.uint64), ind[lo:hi], val[lo:hi]


def test_route_numpy_matches_definition():
    r, c, v = make_coo(50, 70, 400, 0)
    starts = spd.partition_starts(50, 3)
    keys, vals, counts = spd.route_numpy(50, 70, r, c, v, starts)
    assert sum(counts) == len(v)
    pos = 0
    for g, cnt in enumerate(counts):
        lr, lc = spd.unpack_keys(keys[pos:pos + cnt], 70)
        sel = (r >= starts[g]) & (r < starts[g + 1])
        assert np.array_equal(lr + np.uint64(starts[g]), r[sel]) and np.array_equal(lc, c[sel])
        assert vals[pos:pos + cnt].tobytes() == v[sel].tobytes()      # insertion order kept
        pos += cnt


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, ncols = 301, 257
        r, c, v = make_coo(n, ncols, 6000, 5)
        full = orc.compress_from_coo(n, ncols, orc.make_triplets(r, c, v), "row")
        starts = spd.partition_starts(n, world)
        a, b = syn_block(len(v), world, rank)                     # entries block-distributed by index
        keys, vals, counts = spd.route_numpy(n, ncols, r[a:b], c[a:b], v[a:b], starts)
        rk, rv, rc = spd.exchange_routed(dist, torch, torch.from_numpy(keys.astype(np.int64)),
                                         torch.from_numpy(vals), counts)
        lr, lc = spd.unpack_keys(rk.numpy().astype(np.uint64), ncols)
        nloc = starts[rank + 1] - starts[rank]
        got = orc.compress_from_coo(nloc, ncols, orc.make_triplets(lr, lc, rv.numpy()), "row")
        want = shard_of(full, starts, rank)
        ok = (np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
              and got[2].tobytes() == want[2].tobytes() and sum(rc) == len(rv))
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def syn_block(length, world, rank):
    base, extra = divmod(length, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_assembly_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


# ----------------------------------------------------------------------------------- GPU
def _t(a, dt):
    return torch.from_numpy(np.ascontiguousarray(a).astype(dt)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_route_and_packed_assembly_device(dtype, world):
    import spalinalg_b200 as sp
    from spalinalg_b200 import _capi as capi
    ctx = sp.default_context()
    n, ncols = 1000, 777
    r, c, v = make_coo(n, ncols, 40000, 11 + world, dtype)
    full = orc.compress_from_coo(n, ncols, orc.make_triplets(r, c, v), "row")
    starts = spd.partition_starts(n, world)
    shares = []                                    # shares[src][dst] = (keys, vals)
    for src in range(world):
        a, b = syn_block(len(v), world, src)
        keys, vals, counts = spd.route_device(ctx, torch, capi.SPL_CSR, n, ncols, _t(r[a:b], np.int32),
                                              _t(c[a:b], np.int32), _t(v[a:b], dtype), starts)
        wk, wv, wc = spd.route_numpy(n, ncols, r[a:b], c[a:b], v[a:b], starts)
        assert counts == wc
        assert np.array_equal(keys.cpu().numpy().astype(np.uint64), wk)
        assert vals.cpu().numpy().tobytes() == wv.tobytes()
        offs = np.concatenate([[0], np.cumsum(counts)])
        shares.append([(keys[offs[g]:offs[g + 1]], vals[offs[g]:offs[g + 1]]) for g in range(world)])
    for dst in range(world):                       # what the all-to-all delivers: source-rank order
        rk = torch.cat([shares[src][dst][0] for src in range(world)]).contiguous()
        rv = torch.cat([shares[src][dst][1] for src in range(world)]).contiguous()
        nloc = starts[dst + 1] - starts[dst]
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_from_packed_dev(
            ctx._h, capi.SPL_CSR, capi.SPL_F32 if dtype == np.float32 else capi.SPL_F64, nloc, ncols,
            int(rk.numel()), C.c_void_p(rk.data_ptr()), C.c_void_p(rv.data_ptr()), 1, 1, C.byref(h)))
        m = sp.CsrMatrix._wrap(ctx, h)
        want = shard_of(full, starts, dst)
        assert np.array_equal(m.rowptr(), want[0]) and np.array_equal(m.colind(), want[1])
        assert m.values().tobytes() == want[2].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,ncols", [(np.float32, 3000), (np.float64, 3000), (np.float32, 1 << 20)])
def test_packed_assembly_takes_the_hybrid_route(dtype, ncols, monkeypatch):
    """The receive side of the sharded assembly at a size where its list goes through the hybrid
    route (global passes on the high row bits, block kernel with the tail inside), reached at an
    oracle-sized list by lowering the route's threshold; 32-bit packed keys (3 000 columns) and
    64-bit ones (2^20 columns)."""
    import spalinalg_b200 as sp
    from spalinalg_b200 import _capi as capi
    monkeypatch.setenv("SPL_HYBRID_MIN_LEN", "50000")
    ctx = sp.default_context()
    n, world = 40_000, 2
    r, c, v = make_coo(n, ncols, 600_000, 5, dtype)
    full = orc.compress_from_coo(n, ncols, orc.make_triplets(r, c, v), "row")
    starts = spd.partition_starts(n, world)
    parts = []
    for src in range(world):
        a, b = syn_block(len(v), world, src)
        keys, vals, counts = spd.route_device(ctx, torch, capi.SPL_CSR, n, ncols, _t(r[a:b], np.int32),
                                              _t(c[a:b], np.int32), _t(v[a:b], dtype), starts)
        offs = np.concatenate([[0], np.cumsum(counts)])
        parts.append([(keys[offs[g]:offs[g + 1]], vals[offs[g]:offs[g + 1]]) for g in range(world)])
    for dst in range(world):
        rk = torch.cat([parts[src][dst][0] for src in range(world)]).contiguous()
        rv = torch.cat([parts[src][dst][1] for src in range(world)]).contiguous()
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_from_packed_dev(
            ctx._h, capi.SPL_CSR, capi.SPL_F32 if dtype == np.float32 else capi.SPL_F64,
            starts[dst + 1] - starts[dst], ncols, int(rk.numel()), C.c_void_p(rk.data_ptr()),
            C.c_void_p(rv.data_ptr()), 1, 1, C.byref(h)))
        m = sp.CsrMatrix._wrap(ctx, h)
        want = shard_of(full, starts, dst)
        assert np.array_equal(m.rowptr(), want[0]) and np.array_equal(m.colind(), want[1])
        assert m.values().tobytes() == want[2].tobytes()


@pytest.mark.gpu
def test_route_rejects_out_of_bounds():
    import spalinalg_b200 as sp
    from spalinalg_b200 import _capi as capi
    ctx = sp.default_context()
    r = _t([0, 5, 12], np.int32); c = _t([0, 1, 2], np.int32); v = _t([1.0, 2.0, 3.0], np.float64)
    with pytest.raises(sp.Panic):
        spd.route_device(ctx, torch, capi.SPL_CSR, 10, 10, r, c, v, [0, 5, 10])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,world", [(np.float64, 2), (np.float32, 3), (np.float64, 8)])
def test_spmv_peer_gathers_from_owning_slice(dtype, world):
    """x split over `world` separate device buffers; every rank's block must reproduce the
    oracle SpMV rows (random columns: most gathers leave the own slice)."""
    import spalinalg_b200 as sp
    ctx = sp.default_context()
    n = 5000
    rng = np.random.default_rng(3)
    r, c, v = make_coo(n, n, 60000, 21, dtype)
    ptr, ind, val = orc.compress_from_coo(n, n, orc.make_triplets(r, c, v), "row")
    x = rng.standard_normal(n).astype(dtype)
    yw = orc.csr_spmv(n, ptr, ind, val, x)
    scale = orc.csr_spmv(n, ptr, ind, np.abs(val), np.abs(x))
    starts = spd.partition_starts(n, world)
    slices = [_t(x[starts[g]:starts[g + 1]], dtype) for g in range(world)]
    st = (C.c_uint64 * (world + 1))(*starts)
    sl = (C.c_void_p * world)(*[s.data_ptr() for s in slices])
    tol = 1e-12 if dtype == np.float64 else 1e-5
    for g in range(world):
        lp, li, lv = shard_of((ptr, ind, val), starts, g)
        A = sp.CsrMatrix.new(starts[g + 1] - starts[g], n, lp, li, lv)
        y = torch.zeros(starts[g + 1] - starts[g], dtype=slices[0].dtype, device="cuda")
        ctx.check(ctx._lib.spl_spmv_peer(ctx._h, A._h, world, g, C.cast(st, C.c_void_p),
                                         C.cast(sl, C.c_void_p), C.c_void_p(y.data_ptr())))
        ctx.sync()
        got = y.cpu().numpy()
        assert np.all(np.abs(got - yw[starts[g]:starts[g + 1]]) <= tol * scale[starts[g]:starts[g + 1]] + 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,world,n", [(np.float64, 2, 1001), (np.float32, 3, 100003), (np.float64, 8, 70001)])
def test_peer_pull_gathers_every_slice(dtype, world, n):
    """spl_peer_pull with the slices in separate buffers of one process: every rank's full vector
    must end up with all the other slices (odd lengths: unaligned slices take the byte path)."""
    import spalinalg_b200 as sp
    from spalinalg_b200 import _capi as capi
    ctx = sp.default_context()
    x = np.random.default_rng(2).standard_normal(n).astype(dtype)
    starts = spd.partition_starts(n, world)
    slices = [_t(x[starts[g]:starts[g + 1]], dtype) for g in range(world)]
    st = (C.c_uint64 * (world + 1))(*starts)
    sl = (C.c_void_p * world)(*[s.data_ptr() for s in slices])
    for g in range(world):
        full = torch.full((n,), 7.0, dtype=slices[0].dtype, device="cuda")
        torch.cuda.synchronize()
        ctx.check(ctx._lib.spl_peer_pull(ctx._h, capi.SPL_F32 if dtype == np.float32 else capi.SPL_F64, world, g,
                                         C.cast(st, C.c_void_p), C.cast(sl, C.c_void_p), C.c_void_p(full.data_ptr())))
        ctx.sync()
        got = full.cpu().numpy()
        want = x.copy()
        want[starts[g]:starts[g + 1]] = 7.0                      # the own slice is not touched
        assert got.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_peer_barrier_arrival_and_timeout():
    """One rank of a 2-rank barrier on this GPU: the peer's arrival is a flag value written ahead of
    time; without it the bounded spin must give up and report, never hang."""
    import spalinalg_b200 as sp
    ctx = sp.default_context()
    mine = torch.zeros(8, dtype=torch.int32, device="cuda")
    peer = torch.zeros(8, dtype=torch.int32, device="cuda")
    fl = (C.c_void_p * 2)(mine.data_ptr(), peer.data_ptr())
    mine[1] = 1                                                   # rank 1 already arrived at epoch 1
    torch.cuda.synchronize()
    ctx.check(ctx._lib.spl_peer_barrier(ctx._h, 2, 0, C.cast(fl, C.c_void_p), 1, 1000))
    t = C.c_int()
    ctx.check(ctx._lib.spl_peer_barrier_status(ctx._h, C.byref(t)))
    assert t.value == 0 and int(peer[0].item()) == 1              # our arrival reached the peer's block
    ctx.check(ctx._lib.spl_peer_barrier(ctx._h, 2, 0, C.cast(fl, C.c_void_p), 2, 20))   # nobody comes
    with pytest.raises(sp.DeviceError):
        ctx.check(ctx._lib.spl_peer_barrier_status(ctx._h, C.byref(t)))
    assert t.value == 1


@pytest.mark.gpu
def test_sharded_front_end_world1_nccl():
    """The DistCsrMatrix / PeerVector front end with a one-rank NCCL group on this GPU: the same
    host code as the multi-GPU runs (tests/test_multi_gpu.py needs two GPUs), checked against the
    oracle: assembly, peer SpMV with barrier epochs, add, CSR -> CSC."""
    import spalinalg_b200 as sp
    from spalinalg_b200.synthetic_device import device_view
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        ctx = sp.default_context()
        n = 3001
        r, c, v = make_coo(n, n, 50000, 8)
        full = orc.compress_from_coo(n, n, orc.make_triplets(r, c, v), "row")
        D = spd.DistCsrMatrix.from_device_triplets(dist, torch, n, n, _t(r, np.int32), _t(c, np.int32),
                                                   _t(v, np.float64), ctx=ctx)
        assert np.array_equal(D.local.rowptr(), full[0]) and np.array_equal(D.local.colind(), full[1])
        assert D.local.values().tobytes() == full[2].tobytes()
        ex = spd.PeerExchange(ctx, dist)
        for _ in range(2):                                    # second call reuses the receive buffers
            P = spd.DistCsrMatrix.from_device_triplets_peer(dist, torch, n, n, _t(r, np.int32), _t(c, np.int32),
                                                            _t(v, np.float64), ex)
            assert np.array_equal(P.local.rowptr(), full[0]) and np.array_equal(P.local.colind(), full[1])
            assert P.local.values().tobytes() == full[2].tobytes()
        ex.check()
        ex.close()
        x = np.random.default_rng(1).standard_normal(n)
        xv = spd.PeerVector(ctx, dist, n, np.float64)
        device_view(torch, xv.local_ptr, n, torch.float64).copy_(torch.from_numpy(x))
        y = torch.zeros(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        xv.publish()                                          # the written buffer becomes the one products read
        for _ in range(2):
            D.spmv_peer(xv, y.data_ptr())
            xv.barrier()                                      # x unchanged: barrier without the swap
        ctx.sync()
        xv.check()
        yw = orc.csr_spmv(n, *full, x)
        sc = orc.csr_spmv(n, full[0], full[1], np.abs(full[2]), np.abs(x))
        assert np.all(np.abs(y.cpu().numpy() - yw) <= 1e-12 * sc + 1e-300)
        # the windowed product (halo path) with a world of one: the window is the whole of x
        assert D.halo_widths(dist, torch) == (0, 0)
        y.fill_(7.0)
        torch.cuda.synchronize()
        xv.barrier_halo()
        D.spmv_halo(xv, y.data_ptr())
        ctx.sync()
        xv.check()
        assert np.all(np.abs(y.cpu().numpy() - yw) <= 1e-12 * sc + 1e-300)
        # the fused gather kernel with a world of one: nothing to copy, the own block only
        D.prepare_gather(torch)
        xf = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
        y.fill_(7.0)
        torch.cuda.synchronize()
        D.spmv_gather(xv, xf.data_ptr(), y.data_ptr())
        ctx.sync()
        assert np.all(np.abs(y.cpu().numpy() - yw) <= 1e-12 * sc + 1e-300)
        # y_t -> x_{t+1} without a copy: the product writes the next x into the unpublished buffer
        A1 = 0.05 * np.asarray(full[2])
        D1 = spd.DistCsrMatrix(type(D.local).new(n, n, full[0], full[1], A1, ctx=ctx), D.starts, 0, n, n)
        ref = x
        for _ in range(6):
            D1.spmv_peer(xv, xv.local_ptr)
            xv.publish()
            ref = orc.csr_spmv(n, full[0], full[1], A1, ref)
        ctx.sync()
        xv.check()
        got = device_view(torch, xv.published_ptr, n, torch.float64).cpu().numpy()
        scale = np.abs(ref).max() + 1e-300
        assert np.all(np.abs(got - ref) <= 1e-10 * scale)
        S = D + D
        w = orc.addsub(0, n, n, full, full)
        assert np.array_equal(S.local.colind(), w[1]) and S.local.values().tobytes() == w[2].tobytes()
        wc = orc.recompress(n, n, *full)
        ex2 = spd.PeerExchange(ctx, dist)
        for T in (D.to_csc(dist, torch), D.to_csc(dist, torch, exchange=ex2)):
            assert np.array_equal(T.local.colptr(), wc[0]) and np.array_equal(T.local.rowind(), wc[1])
            assert T.local.values().tobytes() == wc[2].tobytes()
        yt, rst = T.spmv(dist, torch, torch.from_numpy(x).cuda())          # column-sharded y = A x (world 1)
        assert rst == [0, n] and np.all(np.abs(yt.cpu().numpy() - yw) <= 1e-12 * sc + 1e-300)
        ex2.close()
        xv.close(dist)
    finally:
        dist.destroy_process_group()
