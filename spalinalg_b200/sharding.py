"""Row sharding of the hot path across the GPUs of one box.

One process per GPU; rank g owns the contiguous row block [r_g, r_{g+1}).  Local CSR blocks keep
GLOBAL column indices, x lives in a global-length buffer on every rank of which the rank owns
(and is the only writer of) slice [r_g, r_{g+1}).  The only data-path exchange of SpMV is x:
  * halo      — each rank sends its first/last `halo` owned elements to its neighbours
                (banded / stencil matrices whose column footprint leaves the block by <= halo);
  * allgather — everybody gets everything (general matrices).
Add/sub/neg need no exchange when the operands share the partition.  torch.distributed is the
plumbing (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def row_partition(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Equal-row contiguous blocks; first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def column_halo(col_min: int, col_max: int, r0: int, r1: int) -> int:
    """Halo width needed by a block whose column indices span [col_min, col_max]."""
    return max(0, r0 - col_min, col_max - (r1 - 1))


def exchange_halo(dist, x_full, r0: int, r1: int, halo: int, rank: int, world: int):
    """Fill x_full[r0-halo:r0] and x_full[r1:r1+halo] from the neighbouring ranks."""
    if world == 1 or halo == 0:
        return
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, x_full[r0:r0 + halo], rank - 1))
        ops.append(dist.P2POp(dist.irecv, x_full[r0 - halo:r0], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, x_full[r1 - halo:r1], rank + 1))
        ops.append(dist.P2POp(dist.irecv, x_full[r1:r1 + halo], rank + 1))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def exchange_allgather(dist, x_full, r0: int, r1: int, world: int, sizes_equal: bool):
    """All ranks receive every owned slice of x."""
    if world == 1:
        return
    if sizes_equal:
        dist.all_gather_into_tensor(x_full, x_full[r0:r1])
    else:
        n = x_full.numel()       # ragged blocks: one broadcast per owner
        for g in range(world):
            a, b = row_partition(n, world, g)
            dist.broadcast(x_full[a:b], src=g)
