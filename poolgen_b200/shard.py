"""Host-side glue of the multi-GPU path when it is launched as one process per GPU.  Everything that
touches data is behind the C ABI: contiguous shard ranges come from `pg_shard_range` -- the same partition the reference
uses for its reader threads (contiguous byte ranges, src/base/helpers.rs:74-91, src/base/sync.rs:917-939), so rank order
= file order -- and the one exchange step of the kinship path, the sum of the per-rank partial Gram matrices
(src/gwas/ols.rs:295 over a column-sharded G), is `pg_kin_allreduce` on the library's own NCCL communicator.  What is
left here is the rendezvous: rank 0 asks the library for the communicator id and ships its 128 bytes to the other ranks
over whatever `torch.distributed` backend the launcher set up (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

from . import capi


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of `rank`: contiguous, sizes differ by at most one, earlier ranks take the larger shards."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"shard_range(total={total}, rank={rank}, world={world})")
    return capi.shard_range(total, rank, world)


def shard_sizes(total: int, world: int) -> list[int]:
    return [e - b for b, e in (shard_range(total, r, world) for r in range(world))]


def broadcast_comm_id(dist=None, make_id=None) -> bytes:
    """rank 0 obtains the communicator id (`make_id`, default pg_comm_unique_id) and every rank receives its bytes"""
    make_id = make_id or capi.Comm.unique_id
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return make_id()
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != capi.COMM_ID_BYTES:
        raise capi.PgError("communicator id broadcast failed")
    return bytes(uid)


def make_comm(ctx, dist=None) -> "capi.Comm":
    """the library's NCCL communicator over the ranks of the running torch.distributed job (one process per GPU)"""
    uid = broadcast_comm_id(dist)
    if dist is None or not dist.is_initialized():
        return capi.Comm.init_rank(ctx, uid, 0, 1)
    return capi.Comm.init_rank(ctx, uid, dist.get_rank(), dist.get_world_size())


def ols_with_covariate_sharded(comm, kin, phen, variance_explained: float = 0.75):
    """ols_with_covariate (src/gwas/ols.rs:278-436) over a column-sharded G: `kin` holds this rank's columns.
    Returns (n_eigenvecs, total columns, beta, var, pval) -- the records of THIS rank's columns, [k, columns]."""
    kin.gram()
    P_total = comm.kin_allreduce([kin])
    m = kin.eig_select(0, variance_explained)  # 0 = the column count the all-reduce summed
    beta, var, pval = kin.covar_scan(phen)
    return m, P_total, beta, var, pval


def gather_in_rank_order(local: np.ndarray, dist=None) -> np.ndarray | None:
    """result slabs concatenated in rank order on rank 0 (= file order, sync.rs:941-967); None elsewhere"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    return np.concatenate(out) if dist.get_rank() == 0 else None
