"""The multi-GPU side of the C ABI on hardware: the library's NCCL communicator and the one exchange step
of ols_iter_with_kinship (pg_kin_allreduce: sum of the per-GPU partial Gram matrices, src/gwas/ols.rs:295 over column
shards).  The one-rank communicator runs on any box; the two-device tests are skipped when fewer than two GPUs are
visible (they ran under `gpurun --gpus 2`, profiles/multigpu_r2.txt)."""
import os
import socket

import numpy as np
import pytest

import poolgen_b200 as pb
from poolgen_b200 import shard

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _problem(n=96, P=3001, k=2, seed=11):
    rng = np.random.default_rng(seed)
    base = rng.random((1, n))
    G = np.clip(0.5 * base + 0.5 * rng.random((P, n)), 0.0, 1.0)   # frequency-like columns with a common component
    phen = rng.standard_normal((n, k))
    return G, phen


def _one_rank(ctx, G, phen, thr):
    kin = pb.Kinship(ctx, G.shape[1], G.shape[0])
    kin.append_columns(G)
    kin.gram()
    K = kin.partial_get()
    m = kin.eig_select(G.shape[0], thr)
    rec = kin.covar_scan(phen)
    kin.close()
    return K, m, rec


def _threshold_for(G, m_target):
    w = np.linalg.eigvalsh(G.T @ G / G.shape[0])[::-1]
    share = np.cumsum(w / w.sum())
    return 0.5 * (share[m_target - 1] + share[m_target]) if m_target > 0 else 0.5 * share[0]


def test_one_rank_communicator_is_the_identity(ctx):
    """ncclCommInitRank with one rank: the all-reduce leaves the matrix as it is and reports the resident columns"""
    G, phen = _problem(n=40, P=500)
    comm = shard.make_comm(ctx)
    assert comm.info() == {"world": 1, "n_local": 1, "first_rank": 0}
    kin = pb.Kinship(ctx, 40, 500)
    kin.append_columns(G)
    kin.gram()
    K = kin.partial_get()
    total, ms = comm.kin_allreduce([kin], timed=True)
    assert total == 500 and ms >= 0.0
    assert np.array_equal(kin.partial_get(), K)
    m = kin.eig_select(0, 0.75)            # 0 = the count the all-reduce summed
    rec = kin.covar_scan(phen)
    kin.close()
    comm.close()
    K1, m1, rec1 = _one_rank(ctx, G, phen, 0.75)
    assert m == m1 and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(rec, rec1))


@pytest.mark.parametrize("m_target", [0, 4])
def test_one_process_two_devices(m_target):
    """pg_init_multi: ONE process drives two GPUs through the C ABI (what the Rust CLI does): column shards, partial
    Gram matrices, pg_kin_allreduce, one eigen step shared with pg_kin_copy_covariates, the covariate scan per shard --
    the records equal the one-GPU records"""
    if _n_devices() < 2:
        pytest.skip("needs two GPUs")
    G, phen = _problem()
    P, n = G.shape
    thr = _threshold_for(G, m_target)
    comm = pb.Comm.init_multi([0, 1])
    try:
        assert comm.info() == {"world": 2, "n_local": 2, "first_rank": 0}
        ranges = [pb.shard_range(P, r, 2) for r in range(2)]
        kins = []
        for c, (b, e) in zip(comm.contexts, ranges):
            kin = pb.Kinship(c, n, e - b)
            kin.append_columns(G[b:e])
            kin.gram()
            kins.append(kin)
        total = comm.kin_allreduce(kins)
        assert total == P
        K0, K1 = kins[0].partial_get(), kins[1].partial_get()
        assert np.array_equal(K0, K1)                       # every rank holds the same sum
        Kw, mw, recw = _one_rank(comm.contexts[0], G, phen, thr)
        assert np.allclose(K0, Kw, rtol=1e-13, atol=0.0)
        m = kins[0].eig_select(0, thr)
        kins[1].copy_covariates_from(kins[0])
        assert m == mw == m_target
        recs = [kin.covar_scan(phen) for kin in kins]
        for kin in kins:
            kin.close()
        for j, name in enumerate(("beta", "var", "pval")):
            got = np.concatenate([r[j] for r in recs], axis=1)
            tol = RTOL if j < 2 else 1e-6
            scale = np.abs(recw[j]) if j != 0 else np.maximum(np.abs(recw[0]), np.sqrt(recw[1]))
            assert got.shape == recw[j].shape and np.all(np.abs(got - recw[j]) <= tol * scale + 2.3e-16), name
    finally:
        comm.close()
        for c in comm.contexts:
            c.close()


def _rank_worker(rank, world, port, out_dir, m_target):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # rendezvous only: the data path is the library's NCCL
    try:
        G, phen = _problem()
        P, n = G.shape
        thr = _threshold_for(G, m_target)
        ctx = pb.Context(rank)
        comm = shard.make_comm(ctx, dist)
        b, e = shard.shard_range(P, rank, world)
        kin = pb.Kinship(ctx, n, e - b)
        kin.append_columns(G[b:e])
        m, P_total, beta, var, pval = shard.ols_with_covariate_sharded(comm, kin, phen, thr)
        kin.close()
        assert P_total == P
        stacked = np.concatenate([beta, var, pval], axis=0).T.copy()      # [columns, 3k]
        got = shard.gather_in_rank_order(stacked, dist)
        if rank == 0:
            Kw, mw, recw = _one_rank(ctx, G, phen, thr)
            k = phen.shape[1]
            assert m == mw == m_target
            gb, gv, gp = got[:, :k].T, got[:, k:2 * k].T, got[:, 2 * k:].T
            assert np.all(np.abs(gb - recw[0]) <= RTOL * np.maximum(np.abs(recw[0]), np.sqrt(recw[1])))
            assert np.all(np.abs(gv - recw[1]) <= RTOL * np.abs(recw[1]))
            assert np.all(np.abs(gp - recw[2]) <= 1e-6 * np.abs(recw[2]) + 2.3e-16)
        comm.close()
        ctx.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("m_target", [0, 4])
def test_two_ranks_one_process_per_gpu(tmp_path, m_target):
    """one process per GPU (the torchrun layout bench.py uses): rank 0 obtains the communicator id from the library, the
    ranks join with pg_comm_init_rank, and the sharded ols_iter_with_kinship records equal the one-rank records"""
    if _n_devices() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_rank_worker, args=(2, port, str(tmp_path), m_target), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
