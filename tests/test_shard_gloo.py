"""world_size-2 gloo tests (CPU) of the host side of the multi-GPU path: the shard ranges of the C ABI
(pg_shard_range), the rendezvous of the library's NCCL communicator id (rank 0 -> every rank) and rank-ordered
gathering.  No CUDA is touched: the exchange step itself (pg_kin_allreduce) is checked on hardware by
tests/test_multigpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poolgen_b200 import capi, shard  # noqa: E402


def test_shard_range_partitions():
    for total in (0, 1, 7, 10_000_000, 1_250_001):
        for world in (1, 2, 3, 8):
            rs = [shard.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == shard.shard_sizes(total, world)
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, P = 12, 101
        rng = np.random.default_rng(5)
        G = rng.random((P, n))                      # every rank builds the same matrix, owns one column shard
        b, e = shard.shard_range(P, rank, world)
        # the partial Gram matrices of contiguous column shards sum to the whole (what pg_kin_allreduce relies on)
        partial = G[b:e].T @ G[b:e]
        import torch
        t = torch.from_numpy(partial.copy())
        dist.all_reduce(t)
        assert np.allclose(t.numpy(), G.T @ G, rtol=1e-13)
        # communicator rendezvous: the id rank 0 obtained from the library reaches every rank unchanged
        uid = shard.broadcast_comm_id(dist)
        assert len(uid) == capi.COMM_ID_BYTES
        box = [None] * world
        dist.all_gather_object(box, uid)
        assert all(u == box[0] for u in box)
        got = shard.gather_in_rank_order(np.arange(b, e), dist)
        if rank == 0:
            assert np.array_equal(got, np.arange(P))
        else:
            assert got is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_protocol(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
