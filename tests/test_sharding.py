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
