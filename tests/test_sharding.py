"""N > 1 host logic on CPU: world_size-2 gloo run of the row partition and the x exchange
(halo and all-gather), with the per-shard SpMV done by the oracle (checker only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spalinalg_b200 import sharding, synthetic as syn


def test_row_partition_covers_rows():
    for n, w in ((10, 3), (100000000, 8), (7, 8), (16, 2)):
        parts = [sharding.row_partition(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    assert sharding.column_halo(6, 23, 10, 20) == 4
    assert sharding.column_halo(12, 18, 10, 20) == 0


def _worker(rank, world, port, n, mode, out):
    import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, c, v = syn.banded(n, range(-4, 5))
        ptr, ind, val = syn.csr_from_sorted_triplets(n, r, c, v)
        x = np.sin(np.arange(n) * 1e-3)
        r0, r1 = sharding.row_partition(n, world, rank)
        lo, hi = int(ptr[r0]), int(ptr[r1])
        lptr = (ptr[r0:r1 + 1] - ptr[r0]).astype(np.uint64)
        lind, lval = ind[lo:hi], val[lo:hi]
        halo = sharding.column_halo(int(lind.min()), int(lind.max()), r0, r1)
        assert halo <= 4
        x_full = torch.full((n,), float("nan"), dtype=torch.float64)
        x_full[r0:r1] = torch.from_numpy(x[r0:r1])            # each rank only knows its own slice
        if mode == "halo":
            sharding.exchange_halo(dist, x_full, r0, r1, 4, rank, world)
        else:
            sharding.exchange_allgather(dist, x_full, r0, r1, world, n % world == 0)
        y_loc = orc.csr_spmv(r1 - r0, lptr, lind, lval, x_full.numpy())
        y_ref = orc.csr_spmv(n, ptr, ind, val, x)[r0:r1]
        ok = bool(np.array_equal(y_loc, y_ref)) and not np.isnan(y_loc).any()
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode,n", [("halo", 1000), ("allgather", 1000), ("allgather", 1001)])
def test_sharded_spmv_world2_gloo(mode, n):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, mode, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


def test_balanced_starts_follow_the_entries():
    """nnz-balanced row partition: contiguous, strictly increasing, every rank gets about
    total / world entries on a power-law row distribution; uniform rows reduce to equal blocks."""
    from spalinalg_b200 import dist as spd
    rng = np.random.default_rng(0)
    counts = (rng.pareto(1.2, 4096) * 100).astype(np.int64)
    rows_per_bin, n = 16, 4096 * 16
    for world in (2, 3, 8):
        st = spd.balanced_starts(counts, rows_per_bin, n, world)
        assert st[0] == 0 and st[-1] == n and all(b > a for a, b in zip(st, st[1:]))
        cum = np.concatenate([[0], np.cumsum(counts + rows_per_bin)])
        share = [cum[st[g + 1] // rows_per_bin] - cum[st[g] // rows_per_bin] for g in range(world)]
        assert max(share) <= 1.15 * cum[-1] / world
    assert spd.balanced_starts([7] * 8, 10, 80, 4) == [0, 20, 40, 60, 80]
    # more ranks than non-empty bins: still a valid partition
    st = spd.balanced_starts([5, 0, 0, 0], 4, 16, 4)
    assert st[0] == 0 and st[-1] == 16 and all(b > a for a, b in zip(st, st[1:]))


def test_gather_groups_block_the_shard_by_ring_offset(monkeypatch):
    """Blocking of a general shard for the fused all-gather + SpMV (spl_spmv_gather_fused): block 0 is the
    own slice alone, the peers follow in ring order, one block per rank unless SPL_GATHER_GROUPS groups them."""
    from spalinalg_b200.dist import DistCsrMatrix as D
    monkeypatch.delenv("SPL_GATHER_GROUPS", raising=False)
    for world in (1, 2, 3, 8):
        first = D.gather_groups(world)
        assert first == list(range(world + 1))
    monkeypatch.setenv("SPL_GATHER_GROUPS", "1,1,2,4")
    assert D.gather_groups(8) == [0, 1, 2, 4, 8]
    for bad, world in (("2,6", 8), ("1,1,2", 8), ("1,0,7", 8)):
        monkeypatch.setenv("SPL_GATHER_GROUPS", bad)
        with pytest.raises(ValueError):
            D.gather_groups(world)


@pytest.mark.parametrize("world,first", [(1, [0, 1]), (2, [0, 1, 2]), (3, [0, 1, 2, 3]), (5, [0, 1, 3, 5]), (8, [0, 1, 2, 4, 8])])
def test_gather_block_layout_is_the_shard_regrouped_by_column_owner(world, first):
    """The blocked form spl_spmv_gather_fused takes (dist.gather_block_layout, the body of prepare_gather) on CPU
    tensors: every entry of the shard appears once, in the block of its column's owner (ring order from the own
    rank), rows and columns in order inside a block; the pointer arrays are padded as the kernel's 16-byte bulk
    copies need; tile_entries_max bounds every tile; and the kernel's walk over the layout (block by block, row
    by row) gives the shard's product."""
    from spalinalg_b200.dist import gather_block_layout, partition_starts
    rng = np.random.default_rng(world)
    n = 2000 + world
    starts = partition_starts(n, world)
    x = rng.standard_normal(n)
    for rank in range(world):
        r0, r1 = starts[rank], starts[rank + 1]
        nloc = r1 - r0
        deg = rng.integers(0, 12, nloc)
        deg[rng.integers(0, nloc, 3)] = 90                                   # a few long rows
        ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        col = np.concatenate([np.sort(rng.choice(n, d, replace=False)) for d in deg]).astype(np.int64)
        val = rng.standard_normal(len(col))
        lay = gather_block_layout(torch, torch.from_numpy(ptr), torch.from_numpy(col.astype(np.uint32).view(np.int32)),
                                  torch.from_numpy(val), starts, rank, first)
        nb, stride = len(first) - 1, lay["stride"]
        bptr, bind, bval = lay["bptr"].numpy().astype(np.int64), lay["bind"].numpy().view(np.uint32).astype(np.int64), lay["bval"].numpy()
        assert bptr.shape == (nb, stride) and stride % 4 == 0 and stride >= nloc + 1 + 3
        assert len(bind) == len(col) + 4 and len(bval) == len(col) + 4
        assert bptr[0, 0] == 0 and bptr[-1, nloc] == len(col)
        assert np.all(bptr[:, nloc + 1:] == bptr[:, nloc:nloc + 1])        # padding repeats the end position
        assert np.all(bptr[1:, 0] == bptr[:-1, nloc])                      # blocks follow each other
        owner = np.searchsorted(np.asarray(starts[1:-1]), bind[:len(col)], side="right")
        y = np.zeros(nloc)
        seen = 0
        for b in range(nb):
            lo, hi = bptr[b, 0], bptr[b, nloc]
            off = (owner[lo:hi] - rank) % world
            assert np.all((off >= first[b]) & (off < first[b + 1]))        # the block holds its owners' columns only
            for r in range(nloc):
                a, e = bptr[b, r], bptr[b, r + 1]
                assert np.all(np.diff(bind[a:e]) > 0)                      # ascending columns inside (block, row)
                y[r] += float(np.dot(bval[a:e], x[bind[a:e]]))
                seen += e - a
            for w, cap in zip((64, 128, 256, 512, 1024), lay["caps"]):
                for r in range(0, nloc, 32):
                    assert bptr[b, min(r + w, nloc)] - bptr[b, r] <= cap
        assert seen == len(col)
        want = np.array([np.dot(val[ptr[r]:ptr[r + 1]], x[col[ptr[r]:ptr[r + 1]]]) for r in range(nloc)])
        assert np.allclose(y, want, rtol=1e-12, atol=1e-12)
