"""Host-side multi-GPU logic of the scan (one process per GPU): contiguous locus / column ranges per
rank -- the same partition the reference uses for its reader threads (contiguous byte ranges, src/base/helpers.rs:74-91,
src/base/sync.rs:917-939), so rank order = file order -- and the one exchange step of the kinship path, the sum of the
per-rank partial Gram matrices (src/gwas/ols.rs:295 over a column-sharded G).  Works with any torch.distributed backend
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of `rank`: contiguous, sizes differ by at most one, earlier ranks take the larger shards."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"shard_range(total={total}, rank={rank}, world={world})")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


class _DevArray:
    """a raw device pointer as a CUDA array (for torch.as_tensor)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def allreduce_partial_gram(kin, dist=None, device_tensor: bool = True):
    """Sum the per-rank partial Gram matrices in place (every rank ends with the total).  With NCCL the library's own
    device buffer is reduced without a copy; with a host backend (gloo) the matrix takes a round trip through numpy."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    import torch
    if device_tensor and dist.get_backend() == "nccl":
        ptr, n = kin.partial_device_ptr()
        t = torch.as_tensor(_DevArray(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    else:
        K = kin.partial_get()
        t = torch.from_numpy(K)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        kin.partial_set(t.numpy())


def total_columns(local_columns: int, dist=None) -> int:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(local_columns)
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(local_columns)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_in_rank_order(local: np.ndarray, dist=None) -> np.ndarray | None:
    """result slabs concatenated in rank order on rank 0 (= file order, sync.rs:941-967); None elsewhere"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    return np.concatenate(out) if dist.get_rank() == 0 else None
