"""GPU parity tests of the ols_iter_with_kinship path (pg_kin_*): column loader, DMMA Gram matrix, PC selection and
the covariate scan against the CPU oracle (oracle/pgo.py: load_columns, ols_with_covariate)."""
import numpy as np
import pytest

import poolgen_b200 as pb
from oracle import pgo
from tests import helpers as H

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # beta, var
PTOL = 1e-6   # p-values (absolute floor 2.3e-16, tests/helpers.py)


def _arbiter(G, cov, phen, c, j):
    """QR least squares of column c / phenotype j (accurate to cond(X) eps, unlike the normal equations the reference
    and the oracle invert): (beta, var) of the allele coefficient."""
    n = G.shape[1]
    x = np.ones((n, 2 + cov.shape[1]))
    x[:, 1:1 + cov.shape[1]] = cov
    x[:, -1] = G[c]
    q, r = np.linalg.qr(x)
    b = np.linalg.solve(r, q.T @ phen[:, j])
    e = phen[:, j] - x @ b
    rinv = np.linalg.inv(r)
    return b[-1], (e @ e) / (n - x.shape[1]) * (rinv[-1] @ rinv[-1])


def _cmp_records(dev, orc, label, arb=None):
    """dev / orc: (beta, var, pval) as [k, P].  Entries outside the tolerance are arbitrated: the device
    passes if it is at least as close to the QR solution as 4x the oracle's own error."""
    beta, var, pval = dev
    ob, ov, op = orc
    assert beta.shape == ob.shape, (beta.shape, ob.shape)
    nan_o = np.isnan(ob)
    assert (np.isnan(beta) == nan_o).all(), label
    ok = ~nan_o
    se = np.sqrt(np.where(ok, ov, 1.0))
    bad_b = ok & ~(np.abs(beta - ob) <= RTOL * np.maximum(np.abs(ob), se))
    bad_v = ok & ~(np.abs(var - ov) <= 4 * RTOL * np.abs(ov))
    bad_p = ok & ~(np.abs(pval - op) <= PTOL * np.abs(op) + H.P_FLOOR)
    bad = bad_b | bad_v | bad_p
    if bad.any():
        assert arb is not None, (label, int(bad.sum()))
        G, cov, phen = arb
        for j, c in zip(*np.nonzero(bad)):
            xb, xv = _arbiter(G, cov, phen, c, j)
            eb_o, eb_d = abs(ob[j, c] - xb), abs(beta[j, c] - xb)
            ev_o, ev_d = abs(ov[j, c] - xv), abs(var[j, c] - xv)
            assert eb_d <= max(4 * eb_o, RTOL * max(abs(xb), np.sqrt(xv))), (label, j, c, beta[j, c], ob[j, c], xb)
            assert ev_d <= max(4 * ev_o, 4 * RTOL * abs(xv)), (label, j, c, var[j, c], ov[j, c], xv)
    return int(bad.sum())


@pytest.mark.parametrize("keep_p_minus_1", [False, True])
@pytest.mark.parametrize("fkw", [dict(), dict(min_coverage_depth=10, min_allele_frequency=0.01),
                                 dict(min_allele_frequency=0.05)])
def test_loader_c1_bit_exact(ctx, fkw, keep_p_minus_1):
    """LoadAll over the reference's tests/test.sync: same columns, same labels, same bits."""
    c1 = H.load_c1()
    fs = pb.FilterStats(pool_sizes=c1["pool_sizes"], **fkw)
    counts = c1["counts"]  # [L, 6, n]
    L = counts.shape[0]
    kin = pb.Kinship(ctx, 5, 5 * L)
    loc, alle = kin.append_counts(counts, c1["codes"], fs, keep_p_minus_1)
    G = kin.get_columns(0, kin.columns)
    kin.close()
    ocols, olabels = pgo.load_columns(counts.transpose(0, 2, 1).astype(np.uint64), c1["codes"], H.oracle_fs(fs),
                                      keep_p_minus_1)
    assert len(olabels) == G.shape[0]
    assert [l for l, _ in olabels] == list(loc)
    assert [a for _, a in olabels] == list(alle)
    assert np.array_equal(G, ocols, equal_nan=True)


def test_loader_six_columns_keep_ns(ctx):
    """--keep-ns with MAF 0 and keep_p_minus_1 off: a locus can emit all SIX columns (the selection word holds six
    column indices next to the kept set; the label arrays hold six entries per locus)"""
    rng = np.random.default_rng(66)
    n, L = 7, 400
    counts = rng.integers(1, 40, size=(L, 6, n)).astype(np.uint32)
    counts[::5, 4] = 0          # some loci without N reads
    counts[::7, 2:4] = 0        # some with two empty columns
    codes = np.arange(6, dtype=np.uint8)
    for keep_p_minus_1 in (False, True):
        fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), remove_ns=False, min_allele_frequency=0.0)
        kin = pb.Kinship(ctx, n, 6 * L)
        loc, alle = kin.append_counts(counts, codes, fs, keep_p_minus_1)
        G = kin.get_columns(0, kin.columns)
        kin.close()
        ocols, olabels = pgo.load_columns(counts.transpose(0, 2, 1).astype(np.uint64), codes, H.oracle_fs(fs), keep_p_minus_1)
        assert len(olabels) == G.shape[0] and (np.bincount(loc).max() == (5 if keep_p_minus_1 else 6))
        assert [l for l, _ in olabels] == list(loc)
        assert [a for _, a in olabels] == list(alle)
        assert np.array_equal(G, ocols, equal_nan=True)


@pytest.mark.parametrize("n,P", [(40, 333), (130, 1000), (257, 64)])
def test_gram_matches_numpy(ctx, n, P):
    rng = np.random.default_rng(n * 1000 + P)
    G = rng.random((P, n))
    kin = pb.Kinship(ctx, n, P)
    kin.append_columns(G)
    kin.gram()
    K = kin.partial_get()
    kin.gram()
    K2 = kin.partial_get()
    kin.close()
    ref = G.T @ G
    assert np.allclose(K, ref, rtol=1e-12, atol=1e-12 * P)
    assert np.array_equal(K, K.T)
    assert np.array_equal(K, K2)  # fixed summation order


@pytest.mark.parametrize("n,L,k,thr", [(24, 150, 2, 0.75), (60, 200, 1, 0.999), (100, 120, 3, 0.9999)])
def test_ols_with_covariate_synthetic(ctx, n, L, k, thr):
    seed = 0x5EED0004 + n
    counts = pb.synth_counts_host(seed, 0, L, n, 4)
    codes = np.arange(4, dtype=np.uint8)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    phen = pb.synth_phen_host(seed, n, k)
    kin = pb.Kinship(ctx, n, 4 * L)
    kin.append_counts(counts, codes, fs)
    P = kin.columns
    G = kin.get_columns(0, P)
    kin.gram()
    m = kin.eig_select(P, thr)
    ev = kin.eigvals(n)
    beta, var, pval = kin.covar_scan(phen)
    kin.close()
    om, ob, ov, op = pgo.ols_with_covariate(G, phen, thr)
    w, V = np.linalg.eigh((G.T @ G) / P)
    w, V = w[::-1], V[:, ::-1]
    assert np.allclose(ev, w, rtol=1e-9, atol=1e-12 * w[0])
    assert m == om
    n_arb = _cmp_records((beta, var, pval), (ob.T, ov.T, op.T), f"kinship n={n} m={m}", arb=(G, V[:, :m], phen))
    print(f"kinship n={n} P={P} m={m}: {n_arb} of {P * k} records arbitrated")


def test_explicit_covariates_and_c1(ctx):
    """config C1 columns with two hand-made covariates (no eigen step): the Frisch-Waugh form against the oracle's
    normal equations for every column."""
    c1 = H.load_c1()
    fs = pb.FilterStats(pool_sizes=c1["pool_sizes"], min_coverage_depth=10, min_allele_frequency=0.01)
    kin = pb.Kinship(ctx, 5, 5 * c1["counts"].shape[0])
    kin.append_counts(c1["counts"], c1["codes"], fs)
    P = kin.columns
    G = kin.get_columns(0, P)
    cov = np.array([[0.3], [-1.2], [0.8], [2.0], [-0.4]])
    kin.set_covariates(cov)
    beta, var, pval = kin.covar_scan(c1["phen"])
    kin.close()
    k = c1["phen"].shape[1]
    ob, ov, op = (np.full((P, k), np.nan) for _ in range(3))
    for c in range(P):
        x = np.ones((5, 3))
        x[:, 1] = cov[:, 0]
        x[:, 2] = G[c]
        rc, b, v, p, _ = pgo.ols(x, c1["phen"])
        if rc == 0:
            ob[c], ov[c], op[c] = b[2], v[2], p[2]
    good = ~np.isnan(ob[:, 0]) & (np.abs(ob[:, 0]) < 1e8)
    assert good.sum() > 0.5 * P
    _cmp_records((beta[:, good], var[:, good], pval[:, good]), (ob[good].T, ov[good].T, op[good].T), "C1 covariates")


def test_partial_sum_equals_whole(ctx):
    """two column shards: the sum of the partial Gram matrices equals the Gram matrix of the whole (the all-reduce step)"""
    n, P = 64, 900
    rng = np.random.default_rng(7)
    G = rng.random((P, n))
    whole = pb.Kinship(ctx, n, P)
    whole.append_columns(G)
    whole.gram()
    K = whole.partial_get()
    whole.close()
    parts = []
    for sl in (slice(0, 400), slice(400, P)):
        kin = pb.Kinship(ctx, n, P)
        kin.append_columns(G[sl])
        kin.gram()
        parts.append(kin.partial_get())
        kin.close()
    assert np.allclose(parts[0] + parts[1], K, rtol=1e-13, atol=1e-10)


def test_sync2csv_from_the_device_loader(ctx):
    """sync2csv (SaveCsv::write_csv, src/base/sync.rs:1182-1262): device column loader -> pg_sort_loci ->
    pg_format_frequency_rows equals the rows built from the oracle's loader, text for text"""
    from tests.test_writer import _oracle_sync2csv_rows
    c1 = H.load_c1()
    fs = pb.FilterStats(pool_sizes=c1["pool_sizes"], min_coverage_depth=5, min_allele_frequency=0.01)
    counts = c1["counts"]
    L = counts.shape[0]
    names = [str(s) for s in c1["chrom_names"]]
    chroms = [names[i] for i in c1["chrom_idx"]]
    pos = [int(p) for p in c1["pos"]]
    kin = pb.Kinship(ctx, 5, 5 * L)
    loc, alle = kin.append_counts(counts, c1["codes"], fs, True)
    G = kin.get_columns(0, kin.columns)
    kin.close()
    order = pb.sort_loci(c1["pos"], chr_names=names, chr_index=c1["chrom_idx"])
    got = pb.format_frequency_rows(G, loc, alle, c1["pos"], locus_order=order, chr_names=names,
                                   chr_index=c1["chrom_idx"]).decode()
    ocols, olabels = pgo.load_columns(counts.transpose(0, 2, 1).astype(np.uint64), c1["codes"], H.oracle_fs(fs), True)
    assert got == _oracle_sync2csv_rows(ocols, olabels, chroms, pos) and got.count("\n") == len(olabels) > 1000


def test_sync2csv_from_sync_text(ctx):
    """text in -> rows out for sync2csv: chunks of sync text through pg_kin_append_sync_text (device parser + LoadAll),
    labels from the parsed text, rows equal to those of the count path"""
    from tests.test_text_gpu import _sync_text
    c1 = H.load_c1()
    fs = pb.FilterStats(pool_sizes=c1["pool_sizes"], min_coverage_depth=5, min_allele_frequency=0.01)
    counts = c1["counts"][:3000]
    L = counts.shape[0]
    names = [str(s) for s in c1["chrom_names"]]
    chroms = [names[i] for i in c1["chrom_idx"][:L]]
    pos = [int(p) for p in c1["pos"][:L]]
    text = _sync_text(counts, chroms, pos, extra=[(7, "# a comment"), (11, "Chromosome1\tx\tN\t" + "\t".join(["1:1:1:1:0:0"] * 5))])
    kin = pb.Kinship(ctx, 5, 5 * L)
    nl, off, p, loc, alle = kin.append_sync_text(text, fs, L, keep_p_minus_1=False)
    G = kin.get_columns(0, kin.columns)
    kin.reset()
    loc2, alle2 = kin.append_counts(counts, c1["codes"], fs, False)
    G2 = kin.get_columns(0, kin.columns)
    kin.close()
    assert nl == L and list(p) == pos
    assert np.array_equal(loc, loc2) and np.array_equal(alle, alle2) and np.array_equal(G, G2, equal_nan=True)
    order = pb.sort_loci(p, text=text, line_offsets=off)
    got = pb.format_frequency_rows(G, loc, alle, p, locus_order=order, text=text, line_offsets=off)
    order2 = pb.sort_loci(pos, chr_names=names, chr_index=c1["chrom_idx"][:L])
    expect = pb.format_frequency_rows(G2, loc2, alle2, pos, locus_order=order2, chr_names=names, chr_index=c1["chrom_idx"][:L])
    assert got == expect and got.count(b"\n") == len(loc) > 1000


def test_c4_shape_gram_and_covariate_scan(ctx):
    """BASELINE config C4's shape (2,000 pools): 16 x 16 tiles of which 136 are computed, many column slices, the
    mirrored reduction -- the Gram matrix against fp64 numpy, and the covariate scan against the oracle's normal
    equations on sampled columns, at the default threshold (no PC selected on frequency data) and at a
    threshold that selects m = 10 PCs (src/gwas/ols.rs:291-370)."""
    n, L, k = 2000, 26_000, 2
    kin = pb.Kinship(ctx, n, 2 * L)
    kin.synth(0x5EED0004, 0, L)
    P = kin.columns
    assert P == 2 * L
    G = kin.get_columns(0, P)
    kin.gram()
    K = kin.partial_get()
    kin.gram()
    K2 = kin.partial_get()
    ref = G.T @ G
    assert np.array_equal(K, K.T)
    assert np.array_equal(K, K2)                      # fixed summation order
    assert np.allclose(K, ref, rtol=1e-12, atol=0.0)
    phen = pb.synth_phen_host(0x5EED0004, n, k)
    rng = np.random.default_rng(4)
    sample = np.sort(rng.choice(P, size=600, replace=False))
    # default threshold
    m0 = kin.eig_select(P, 0.75)
    ev = kin.eigvals(n)
    beta, var, pval = kin.covar_scan(phen)
    om, ob, ov, op, w, V = pgo.ols_with_covariate(G, phen, 0.75, columns=sample, return_eig=True)
    assert m0 == om == 0
    assert np.allclose(ev, w, rtol=1e-9, atol=1e-12 * w[0])
    _cmp_records((beta[:, sample], var[:, sample], pval[:, sample]), (ob[sample].T, ov[sample].T, op[sample].T),
                 "C4 shape m=0", arb=(G[sample], V[:, :0], phen))
    # a threshold between the cumulative shares of 10 and 11 eigenvalues selects m = 10
    share = np.cumsum(w / w.sum())
    thr = 0.5 * (share[9] + share[10])
    m10 = kin.eig_select(P, thr)
    beta_d, var_d, pval_d = kin.covar_scan(phen)         # with the device's own eigenvectors
    # the covariate scan itself against the oracle, both on the oracle's PCs: the bulk eigenvalues of a 2,000-pool
    # kinship matrix sit 5e-4 apart (relative), so a 1e-13 difference between two correct Gram matrices turns v_10
    # into v_11 by ~1e-10 -- the PC SUBSPACE is only defined to that level (eigen step parity unpinned),
    # and it is checked separately below at the tolerance its conditioning allows
    kin.set_covariates(V[:, :10])
    beta, var, pval = kin.covar_scan(phen)
    kin.close()
    om, ob, ov, op = pgo.ols_with_covariate(G, phen, thr, columns=sample)
    assert m10 == om == 10
    n_arb = _cmp_records((beta[:, sample], var[:, sample], pval[:, sample]), (ob[sample].T, ov[sample].T, op[sample].T),
                         "C4 shape m=10", arb=(G[sample], V[:, :10], phen))
    ok = ~np.isnan(ob[sample].T)
    se = np.sqrt(ov[sample].T[ok])
    assert np.all(np.abs(beta_d[:, sample][ok] - ob[sample].T[ok]) <= 1e-7 * np.maximum(np.abs(ob[sample].T[ok]), se))
    assert np.all(np.abs(pval_d[:, sample][ok] - op[sample].T[ok]) <= 1e-6 * op[sample].T[ok] + H.P_FLOOR)
    print(f"C4 shape: P={P}, m=0 and m=10 records match on {sample.size} sampled columns ({n_arb} arbitrated)")


@pytest.mark.parametrize("n,L,k", [(24, 150, 2), (60, 120, 1)])
def test_threshold_never_reached_takes_the_n_less_than_p_branch(ctx, n, L, k):
    """a threshold the cumulative shares never reach leaves n_eigenvecs = n (src/gwas/ols.rs:300-311): X = [1 | V | g]
    has n + 2 columns for n pools and the reference solves b = X'(XX')^-1 y (ols.rs:67-75); var is rounding noise over
    a negative n - p, so t is NaN and every p-value is forced to 1 (ols.rs:150-151)"""
    seed = 0x5EED0004 + 7 * n
    counts = pb.synth_counts_host(seed, 0, L, n, 4)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    phen = pb.synth_phen_host(seed, n, k)
    kin = pb.Kinship(ctx, n, 4 * L)
    kin.append_counts(counts, np.arange(4, dtype=np.uint8), fs)
    P = kin.columns
    G = kin.get_columns(0, P)
    kin.gram()
    m = kin.eig_select(P, 1.5)
    beta, var, pval = kin.covar_scan(phen)
    kin.close()
    om, ob, ov, op = pgo.ols_with_covariate(G, phen, 1.5)
    assert m == om == n
    ok = ~np.isnan(ob)
    assert ok.mean() > 0.9
    scale = np.maximum(np.abs(ob), np.abs(ob[ok]).mean())
    assert np.all(np.abs(beta.T - ob)[ok] <= 1e-8 * scale[ok])
    assert np.all(pval.T[ok] == 1.0) and np.all(op[ok] == 1.0)


@pytest.mark.parametrize("n,P,m,k", [(64, 333, 3, 1), (301, 1001, 6, 3), (130, 50, 12, 4), (97, 7, 9, 2),
                                     (2000, 260, 10, 1),    # V = 193 KB: two pool passes (1,008 + 992 pools)
                                     (2900, 150, 12, 3),    # 16 vectors: three pool passes, a middle pass that adds AND parks
                                     (1000, 200, 13, 1)])   # 15 vectors: the list kernel's 13..16-vector instantiations
def test_covariate_scan_dmma_kernel(ctx, n, P, m, k):
    """the blocked FP64 contraction of the covariate scan (covar_mma_kernel: 5..16 vectors): pool counts that are not
    multiples of 16, column counts that are not multiples of 16, one and two M tiles, several phenotypes per column,
    a constant column (singular X'X -> NaN like the reference) -- against the oracle's normal equations"""
    rng = np.random.default_rng(1000 * n + m)
    G = np.clip(0.5 + 0.2 * rng.standard_normal((P, n)), 0.0, 1.0)
    G[P // 2] = 1.0
    # nearly constant columns (a rare allele): the centred g'g loses its digits -> the explicit-residual list kernel
    for c in (1, P // 3, P - 2):
        G[c] = 0.999 + 1e-4 * rng.standard_normal(n)
    cov = rng.standard_normal((n, m))
    phen = rng.standard_normal((n, k))
    kin = pb.Kinship(ctx, n, P)
    kin.append_columns(G)
    kin.set_covariates(cov)
    beta, var, pval = kin.covar_scan(phen)
    kin.close()
    ob, ov, op = (np.full((P, k), np.nan) for _ in range(3))
    for c in range(P):
        x = np.ones((n, 2 + m))
        x[:, 1:1 + m] = cov
        x[:, 1 + m] = G[c]
        rc, b, v, p_, _ = pgo.ols(x, phen)
        if rc == 0 and c != P // 2:
            ob[c], ov[c], op[c] = b[1 + m], v[1 + m], p_[1 + m]
    assert np.isnan(beta[:, P // 2]).all()
    _cmp_records((beta, var, pval), (ob.T, ov.T, op.T), f"dmma covar n={n} m={m}", arb=(G, cov, phen))


@pytest.mark.parametrize("n,m,k", [(3, 0, 3), (4, 1, 2), (5, 2, 2), (6, 3, 1), (7, 4, 3), (9, 2, 4)])
def test_covariate_scan_few_pools(ctx, n, m, k):
    """a handful of pools: near-perfect fits are the rule (1 - r^2 of three points piles up at 0) and the columns of a
    deeply covered locus are nearly constant, so y~'y~ - b g~'y~ loses the digits the centred g'g lost times
    y~'y~ / rss -- those columns take the explicit-residual kernel (found by tools/fuzz_parity.py --mode kin: the
    variance was off by 1e-8 .. 1e-6 where the oracle's explicit residuals keep 1e-13)"""
    rng = np.random.default_rng(100 * n + m)
    P = 4000
    base = rng.uniform(0.05, 0.95, (P, 1))
    G = np.clip(base + rng.uniform(1e-4, 0.2, (P, 1)) * rng.standard_normal((P, n)), 0.0, 1.0)
    cov = rng.standard_normal((n, m))
    phen = rng.standard_normal((n, k)) * 5.0 + 20.0
    kin = pb.Kinship(ctx, n, P)
    kin.append_columns(G)
    kin.set_covariates(cov)
    beta, var, pval = kin.covar_scan(phen)
    kin.close()
    ob, ov, op = (np.full((P, k), np.nan) for _ in range(3))
    for c in range(P):
        x = np.ones((n, 2 + m))
        x[:, 1:1 + m] = cov
        x[:, 1 + m] = G[c]
        rc, b, v, p_, _ = pgo.ols(x, phen)
        if rc == 0 and np.ptp(G[c]) > 0:
            ob[c], ov[c], op[c] = b[1 + m], v[1 + m], p_[1 + m]
    flat = np.ptp(G, axis=1) == 0
    beta[:, flat], var[:, flat], pval[:, flat] = np.nan, np.nan, np.nan
    n_arb = _cmp_records((beta, var, pval), (ob.T, ov.T, op.T), f"few pools n={n} m={m}", arb=(G, cov, phen))
    print(f"few pools n={n} m={m} k={k}: {n_arb} of {P * k} records arbitrated")
