"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs."""
import numpy as np
import pytest

import poolgen_b200 as pb
from oracle import pgo
from tests import helpers as H

pytestmark = pytest.mark.gpu

C1_FILTERS = [
    dict(),                                                        # CLI defaults (main.rs:80-93)
    dict(min_coverage_depth=10, min_allele_frequency=0.01),        # the CI flag set (.github/workflows/rust.yml)
    dict(min_allele_frequency=0.05),                               # q lands exactly on the threshold
    dict(min_allele_frequency=0.1),
    dict(min_coverage_depth=0, max_missingness_rate=0.5),          # NaN frequencies survive the depth filter
]


def _fs(pool_sizes, **kw):
    return pb.FilterStats(pool_sizes=pool_sizes, **kw)


@pytest.mark.parametrize("fkw", C1_FILTERS)
@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR])
def test_c1_regression(ctx, kind, fkw):
    c1 = H.load_c1()
    fs = _fs(c1["pool_sizes"], **fkw)
    scan = pb.Scan(ctx, kind, fs, 5, c1["codes"], c1["phen"])
    dev = scan.run_counts(c1["counts"])
    scan.close()
    st = H.compare_regression(kind, c1["counts"], c1["codes"], c1["phen"], fs, dev, label=f"C1 {kind} {fkw}")
    if not fkw:
        assert st["kept"] == 6556  # SURVEY.md 8a F2 (restated filter on tests/test.sync)
    if fkw == dict(min_coverage_depth=10, min_allele_frequency=0.01):
        assert st["kept"] == 2032
    print(st)


@pytest.mark.parametrize("fkw", C1_FILTERS[:4])
@pytest.mark.parametrize("kind", [pb.KIND_CHISQ, pb.KIND_FISHER])
def test_c1_tables(ctx, kind, fkw):
    c1 = H.load_c1()
    fs = _fs(c1["pool_sizes"], **fkw)
    scan = pb.Scan(ctx, kind, fs, 5, c1["codes"])
    dev = scan.run_counts(c1["counts"])
    scan.close()
    print(H.compare_tables(kind, c1["counts"], c1["codes"], fs, dev, label=f"C1 tables {kind} {fkw}"))


SHAPES = [
    # n_pools, n_alleles, k, loci, weighted
    (100, 4, 1, 20000, False),   # C2 shape
    (1000, 4, 3, 1500, False),   # C3 shape
    (37, 4, 2, 4000, True),      # odd pool count, unequal pool sizes
    (129, 5, 4, 1500, False),    # one row past a chunk boundary, D column present
    (257, 6, 5, 600, True),      # N column dropped on the device, k > 4 (two phenotype passes)
    (8, 2, 1, 3000, False),      # biallelic
    (64, 3, 3, 3000, False),
]


@pytest.mark.parametrize("n,A,k,L,weighted", SHAPES)
@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR])
def test_synthetic_regression(ctx, kind, n, A, k, L, weighted):
    seed = 0x5EED0000 + n * 7 + A
    counts = pb.synth_counts_host(seed, 0, L, n, A)
    phen = pb.synth_phen_host(seed, n, k)
    ps = np.ones(n)
    if weighted:
        ps = 10.0 + (np.arange(n) % 7)
    tot = 0.0
    for v in ps:
        tot = tot + v
    ps = np.array([v / tot for v in ps])
    fs = _fs(ps)
    codes = np.arange(A, dtype=np.uint8)
    scan = pb.Scan(ctx, kind, fs, n, codes, phen)
    dev = scan.run_counts(counts)
    scan.close()
    st = H.compare_regression(kind, counts, codes, phen, fs, dev, label=f"synth n={n} A={A} k={k}")
    assert st["ok"] > 0.8 * L * (0.5 if A < 4 else 1.0) or A < 4
    print(st)


def test_device_generator_matches_host(ctx):
    n, A, k, L = 100, 4, 1, 5000
    seed = 0x5EED0002
    phen = pb.synth_phen_host(seed, n, k)
    fs = _fs(np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, np.arange(A, dtype=np.uint8), phen)
    b = scan.batch(L)
    b.synth(seed, 123, L)
    b.run()
    dev_synth = b.fetch()
    counts = pb.synth_counts_host(seed, 123, L, n, A)
    b.upload_counts(counts)
    b.run()
    dev_host = b.fetch()
    b.close()
    scan.close()
    assert (dev_synth.status == dev_host.status).all()
    assert np.array_equal(dev_synth.stats, dev_host.stats, equal_nan=True)


def test_upload_formats_agree(ctx):
    """u32, u16 and u8 counts and the host-built f64 frequency matrix give identical records."""
    n, A, k, L = 100, 4, 2, 3000
    seed = 77
    counts = pb.synth_counts_host(seed, 0, L, n, A)
    phen = pb.synth_phen_host(seed, n, k)
    fs = _fs(np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, np.arange(A, dtype=np.uint8), phen)
    b = scan.batch(L)
    b.upload_counts(counts)
    b.run()
    r32 = b.fetch()
    b.upload_counts(counts.astype(np.uint16))
    b.run()
    r16 = b.fetch()
    assert counts.max() < 256
    b.upload_counts(counts.astype(np.uint8))
    b.run()
    r8 = b.fetch()
    depth = counts.sum(axis=1).astype(np.uint32)
    with np.errstate(invalid="ignore", divide="ignore"):
        freq = counts.astype(np.float64) / depth[:, None, :].astype(np.float64)
    b.upload_freq(freq, depth)
    b.run()
    rf = b.fetch()
    b.close()
    scan.close()
    for r in (r16, r8, rf):
        assert (r.status == r32.status).all()
        assert np.array_equal(r.stats, r32.stats, equal_nan=True)
        assert np.array_equal(r.freq_mean, r32.freq_mean, equal_nan=True)


def test_streaming_matches_batch(ctx):
    n, A, k, L = 64, 4, 1, 9000
    counts = pb.synth_counts_host(5, 0, L, n, A)
    phen = pb.synth_phen_host(5, n, k)
    fs = _fs(np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_CORR, fs, n, np.arange(A, dtype=np.uint8), phen)
    whole = scan.run_counts(counts)
    scan.stream_begin(2048)
    tickets, parts = [], []
    for i, l0 in enumerate(range(0, L, 2048)):
        slab = np.ascontiguousarray(counts[l0:l0 + 2048])
        tickets.append((scan.submit_counts(slab), slab))
        if len(tickets) == 3:
            t, _ = tickets.pop(0)
            parts.append(scan.collect(t))
    for t, _ in tickets:
        parts.append(scan.collect(t))
    scan.close()
    status = np.concatenate([p.status for p in parts])
    stats = np.concatenate([p.stats for p in parts])
    assert (status == whole.status).all()
    assert np.array_equal(stats, whole.stats, equal_nan=True)


def test_empty_and_tiny_batches(ctx):
    n, A = 10, 4
    phen = pb.synth_phen_host(1, n, 1)
    fs = _fs(np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, np.arange(A, dtype=np.uint8), phen)
    r0 = scan.run_counts(np.zeros((0, A, n), dtype=np.uint32))
    assert r0.status.size == 0
    counts = pb.synth_counts_host(9, 0, 1, n, A)
    r1 = scan.run_counts(counts)
    H.compare_regression(pb.KIND_OLS, counts, np.arange(A, dtype=np.uint8), phen, fs, r1, label="single locus")
    # all-zero counts: every pool has no coverage
    z = np.zeros((3, A, n), dtype=np.uint32)
    rz = scan.run_counts(z)
    assert (rz.status == pb.LOCUS_FILTERED).all()
    scan.close()


def test_c5_tables_synthetic(ctx):
    """config C5 shape: 2 pools, count path, depth 20..100 so Fisher's > 34 rescale is exercised."""
    n, A, L = 2, 4, 20000
    counts = pb.synth_counts_host(0x5EED0005, 0, L, n, A)
    fs = _fs(np.full(n, 0.5))
    codes = np.arange(A, dtype=np.uint8)
    for kind in (pb.KIND_CHISQ, pb.KIND_FISHER):
        scan = pb.Scan(ctx, kind, fs, n, codes)
        dev = scan.run_counts(counts)
        scan.close()
        print(H.compare_tables(kind, counts, codes, fs, dev, label=f"C5 {kind}"))


def test_full_size_c2_properties(ctx):
    """C2 at full size (100 pools x 1M loci): sampled loci against the oracle (host replay of the integer
    generator) and a split-invariance property: scanning two halves gives the bits of the whole."""
    n, A, k, L = 100, 4, 1, 1_000_000
    seed = 0x5EED0002
    phen = pb.synth_phen_host(seed, n, k)
    fs = _fs(np.full(n, 1.0 / n))
    codes = np.arange(A, dtype=np.uint8)
    scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, codes, phen)
    b = scan.batch(L)
    b.synth(seed, 0, L)
    b.run()
    whole = b.fetch()
    # sampled windows against the oracle
    for l0 in (0, 333_333, L - 4096):
        counts = pb.synth_counts_host(seed, l0, 4096, n, A)
        sub = pb.ScanResults(whole.status[l0:l0 + 4096], whole.n_out[l0:l0 + 4096], whole.alleles[l0:l0 + 4096],
                             whole.freq_mean[l0:l0 + 4096], whole.stats[l0:l0 + 4096])
        H.compare_regression(pb.KIND_OLS, counts, codes, phen, fs, sub, label=f"C2 window {l0}")
    # split invariance
    half = L // 2
    b.synth(seed, half, half)
    b.run()
    second = b.fetch()
    assert (second.status == whole.status[half:]).all()
    assert np.array_equal(second.stats, whole.stats[half:], equal_nan=True)
    # sanity of the class mix of the generator: ~5 % monomorphic + ~2 % zero-depth loci are filtered
    frac = (whole.status == pb.LOCUS_FILTERED).mean()
    assert 0.04 < frac < 0.12, frac
    b.close()
    scan.close()


MORE_SHAPES = [
    # n_pools, n_alleles, k, loci, weighted
    (300, 3, 2, 1500, False),    # A=3: chunks of 192 pools, two chunks
    (600, 2, 1, 1500, True),     # A=2: chunks of 256 pools, three chunks, unequal pool sizes
    (1000, 6, 3, 400, True),     # N dropped on the device -> 5 columns, 3 phenotypes in one pass, weighted, 8 chunks
    (130, 4, 4, 1500, False),    # one row pair past a chunk boundary, 4 phenotypes
    (20, 4, 2, 4000, False),     # 8 lanes per locus (tiny pool count)
    (128, 4, 1, 3000, False),    # exactly one chunk, 16 lanes per locus
]


@pytest.mark.parametrize("n,A,k,L,weighted", MORE_SHAPES)
@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR])
def test_more_shapes(ctx, kind, n, A, k, L, weighted):
    seed = 0xABCD0000 + n * 11 + A
    counts = pb.synth_counts_host(seed, 0, L, n, A)
    phen = pb.synth_phen_host(seed, n, k)
    ps = (10.0 + (np.arange(n) % 7)) if weighted else np.ones(n)
    tot = 0.0
    for v in ps:
        tot = tot + v
    fs = _fs(np.array([v / tot for v in ps]))
    codes = np.arange(A, dtype=np.uint8)
    scan = pb.Scan(ctx, kind, fs, n, codes, phen)
    dev = scan.run_counts(counts)
    scan.close()
    print(H.compare_regression(kind, counts, codes, phen, fs, dev, label=f"more n={n} A={A} k={k}"))


@pytest.mark.parametrize("n,L", [(100, 6000), (1000, 600)])
@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR])
def test_error_reads_defer_most_loci(ctx, kind, n, L):
    """real pool-seq data: stray reads on a third / fourth allele in a few pools.  Those alleles fail the MAF filter
    but carry reads, so nearly every locus takes the renormalising fix-up path."""
    seed = 0xE44 + n
    counts = pb.synth_counts_host(seed, 0, L, n, 2)  # biallelic A/T
    full = np.zeros((L, 5, n), dtype=np.uint32)       # A T C G D
    full[:, :2] = counts
    rng = np.random.default_rng(seed)
    for col in (2, 3, 4):
        hit = rng.random((L, n)) < (0.02 if col < 4 else 0.005)
        full[:, col] = hit.astype(np.uint32)
    phen = pb.synth_phen_host(seed, n, 2)
    fs = _fs(np.full(n, 1.0 / n), min_allele_frequency=0.001)
    codes = np.array([0, 1, 2, 3, 5], dtype=np.uint8)
    scan = pb.Scan(ctx, kind, fs, n, codes, phen)
    dev = scan.run_counts(full)
    scan.close()
    st = H.compare_regression(kind, full, codes, phen, fs, dev, label=f"error reads n={n}")
    assert st["ok"] > 0.8 * L
    print(st)


@pytest.mark.parametrize("n,L,weighted", [(40, 4000, False), (100, 3000, True), (300, 1200, True), (1000, 400, False)])
@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR])
def test_hinted_renormalisation(ctx, kind, n, L, weighted):
    """the ingest hint lets the streaming kernel renormalise on the fly (8 / 16 / 32 lanes per locus, a partial last
    row chunk at 300 pools, unequal pool sizes); a pool whose reads all sit on removed alleles makes the renormalised
    frequencies NaN (0 / 0 in to_frequencies, src/base/sync.rs:166-192) and has to come out like the reference's."""
    seed = 0x417 + n
    counts = pb.synth_counts_host(seed, 0, L, n, 3)  # A/T/C
    full = np.zeros((L, 5, n), dtype=np.uint32)
    full[:, :3] = counts
    rng = np.random.default_rng(seed)
    full[:, 3] = (rng.random((L, n)) < 0.03).astype(np.uint32) * rng.integers(1, 3, (L, n)).astype(np.uint32)
    full[:, 4] = (rng.random((L, n)) < 0.004).astype(np.uint32)
    for l in range(7, L, 97):   # pool l % n keeps only stray reads on G
        full[l, :, l % n] = 0
        full[l, 3, l % n] = 4
    phen = pb.synth_phen_host(seed, n, 2)
    ps = (5.0 + (np.arange(n) % 5)) if weighted else np.ones(n)
    tot = 0.0
    for v in ps:
        tot = tot + v
    fs = _fs(np.array([v / tot for v in ps]), min_allele_frequency=0.002)
    codes = np.array([0, 1, 2, 3, 5], dtype=np.uint8)
    scan = pb.Scan(ctx, kind, fs, n, codes, phen)
    dev = scan.run_counts(full)
    scan.close()
    st = H.compare_regression(kind, full, codes, phen, fs, dev, label=f"hinted n={n}")
    assert st["ok"] > 0.7 * L
    print(st)


def test_missing_coverage_at_scale(ctx):
    """--min-coverage-depth 0 with pools without coverage at 300 pools: NaN frequencies, exact q, missingness"""
    n, A, k, L = 300, 4, 2, 1500
    seed = 0x5EED0303
    counts = pb.synth_counts_host(seed, 0, L, n, A)   # 2 % of the loci have a pool at depth 0
    phen = pb.synth_phen_host(seed, n, k)
    fs = _fs(np.full(n, 1.0 / n), min_coverage_depth=0, max_missingness_rate=0.25)
    codes = np.arange(A, dtype=np.uint8)
    for kind in (pb.KIND_OLS, pb.KIND_CORR):
        scan = pb.Scan(ctx, kind, fs, n, codes, phen)
        dev = scan.run_counts(counts)
        scan.close()
        print(H.compare_regression(kind, counts, codes, phen, fs, dev, label="missing coverage"))


def test_full_size_c3_shard_properties(ctx):
    """C3 shape at 300,000 loci (10 GB resident): sampled windows against the oracle and split invariance."""
    n, A, k, L = 1000, 4, 3, 300_000
    seed = 0x5EED0003
    phen = pb.synth_phen_host(seed, n, k)
    fs = _fs(np.full(n, 1.0 / n))
    codes = np.arange(A, dtype=np.uint8)
    scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, codes, phen)
    b = scan.batch(L)
    b.synth(seed, 0, L)
    b.run()
    whole = b.fetch()
    for l0 in (0, 123_456, L - 384):
        counts = pb.synth_counts_host(seed, l0, 384, n, A)
        sub = pb.ScanResults(whole.status[l0:l0 + 384], whole.n_out[l0:l0 + 384], whole.alleles[l0:l0 + 384],
                             whole.freq_mean[l0:l0 + 384], whole.stats[l0:l0 + 384])
        H.compare_regression(pb.KIND_OLS, counts, codes, phen, fs, sub, label=f"C3 window {l0}")
    third = L // 3
    b.synth(seed, third, third)
    b.run()
    part = b.fetch()
    assert (part.status == whole.status[third:2 * third]).all()
    assert np.array_equal(part.stats, whole.stats[third:2 * third], equal_nan=True)
    assert np.array_equal(part.freq_mean, whole.freq_mean[third:2 * third], equal_nan=True)
    frac = (whole.status == pb.LOCUS_FILTERED).mean()
    assert 0.04 < frac < 0.12, frac
    b.close()
    scan.close()


@pytest.mark.parametrize("n,A", [(2, 4), (2, 5), (2, 6), (3, 4), (3, 6), (4, 4), (4, 6), (3, 5), (6, 4)])
@pytest.mark.parametrize("kind", [pb.KIND_CHISQ, pb.KIND_FISHER])
def test_small_tables(ctx, kind, n, A):
    """register-resident table kernels (2-4 pools) and the generic kernel (other shapes) against the oracle; the N
    column (code 4) is dropped on the device when present"""
    L = 6000
    counts = pb.synth_counts_host(0x7AB1E + n * 10 + A, 0, L, n, A)
    if A >= 5:  # put a few reads on the N / D columns so that they matter
        rng = np.random.default_rng(n * 100 + A)
        counts[:, 4:] = (rng.random((L, A - 4, n)) < 0.3).astype(np.uint32) * rng.integers(1, 4, (L, A - 4, n), dtype=np.uint32)
    fs = _fs(np.full(n, 1.0 / n), min_allele_frequency=0.01)
    codes = np.arange(A, dtype=np.uint8)
    scan = pb.Scan(ctx, kind, fs, n, codes)
    dev = scan.run_counts(counts)
    scan.close()
    st = H.compare_tables(kind, counts, codes, fs, dev, label=f"tables n={n} A={A}")
    assert st["ok"] > 0.5 * L
    print(st)


def test_many_pools_split_phenotype_passes(ctx):
    """4,000 pools x 3 phenotypes: the resident phenotype vectors of one pass are capped at 96 KB of shared memory, so
    the scan runs 3 phenotypes as passes of two and one"""
    n, A, k, L = 4000, 4, 3, 96
    seed = 0x4000
    counts = pb.synth_counts_host(seed, 0, L, n, A)
    phen = pb.synth_phen_host(seed, n, k)
    fs = _fs(np.full(n, 1.0 / n))
    codes = np.arange(A, dtype=np.uint8)
    for kind in (pb.KIND_OLS, pb.KIND_CORR):
        scan = pb.Scan(ctx, kind, fs, n, codes, phen)
        dev = scan.run_counts(counts)
        scan.close()
        print(H.compare_regression(kind, counts, codes, phen, fs, dev, label="4000 pools"))


@pytest.mark.parametrize("n,L", [(40, 3000), (300, 800)])
def test_pearson_with_missing_phenotypes(ctx, n, L):
    """NA phenotype values (src/base/phen.rs:68-75): pearson_corr drops the pairs per phenotype
    (correlation_test.rs:21-31); ols_iter panics in the reference, so the scan refuses to open"""
    seed = 0x9A9 + n
    counts = pb.synth_counts_host(seed, 0, L, n, 4)
    phen = pb.synth_phen_host(seed, n, 3)
    phen[3, 0] = np.nan
    phen[7, 0] = np.nan
    phen[n - 1, 2] = np.nan
    fs = _fs(np.full(n, 1.0 / n))
    codes = np.arange(4, dtype=np.uint8)
    scan = pb.Scan(ctx, pb.KIND_CORR, fs, n, codes, phen)
    dev = scan.run_counts(counts)
    scan.close()
    st = H.compare_regression(pb.KIND_CORR, counts, codes, phen, fs, dev, label=f"missing phenotypes n={n}")
    assert st["ok"] > 0.8 * L
    print(st)
    with pytest.raises(pb.PgError):
        pb.Scan(ctx, pb.KIND_OLS, fs, n, codes, phen)


def _sparse_counts(rng, L, A, n, mean_total, cover):
    """count tables for the many-pool Fisher test: a sparse background, a few heavy pools per locus (their rows survive
    the rescale to a total of 34), optionally one read per pool so that every pool has coverage"""
    lam = mean_total / (n * min(A, 4))
    c = np.zeros((L, A, n), dtype=np.uint32)
    c[:, :min(A, 4)] = rng.poisson(lam, size=(L, min(A, 4), n))
    heavy = rng.integers(0, n, size=(L, 3))
    for l in range(0, L, 2):
        for h in heavy[l]:
            c[l, :min(A, 4), h] += rng.integers(0, 12 * max(1, n // 10), size=min(A, 4)).astype(np.uint32)
    if cover:
        c[:, 0] += 1
    return c


@pytest.mark.parametrize("n,A,L,mean_total,kw", [
    (17, 4, 1500, 60, {}), (40, 6, 800, 30, {}), (100, 4, 300, 20, {}), (300, 4, 40, 200, {}),
    (40, 4, 800, 25, dict(min_coverage_depth=0, max_missingness_rate=1.0)),
    (100, 5, 300, 30, dict(min_coverage_depth=0, max_missingness_rate=1.0))])
def test_fisher_many_pools(ctx, n, A, L, mean_total, kw):
    """tables::fisher beyond 16 pools (src/tables/fisher_exact_test.rs:32-130 has no limit): the table is rescaled to a
    total of at most 34, so only its non-zero rows are enumerated; rescaled (total > 34) and raw tables (pools without
    coverage allowed), pools whose rescaled row is empty, the last pool empty or not, all-zero rescaled tables"""
    rng = np.random.default_rng(0xF154 + n + A)
    full = _sparse_counts(rng, L, A, n, mean_total, cover=not kw)
    full[::3, :, n - 1] = 0
    full[::3, 0, n - 1] = 1                                # the last row stays non-zero only through allele A
    fs = _fs(np.full(n, 1.0 / n), min_allele_frequency=0.0, **kw)
    codes = np.arange(A, dtype=np.uint8)
    scan = pb.Scan(ctx, pb.KIND_FISHER, fs, n, codes)
    dev = scan.run_counts(full)
    scan.close()
    st = H.compare_tables(pb.KIND_FISHER, full, codes, fs, dev, label=f"fisher n={n}")
    assert st["ok"] > 0.5 * L
    ok = dev.status == pb.LOCUS_OK
    assert len(np.unique(dev.stats[ok, 0, 0, 3])) > 10     # not only all-zero tables
    print(st)


def test_wide_table_kernels_on_small_tables():
    """PG_TABLES_WIDE=1 (read once per process) sends tables of up to 16 pools through the many-pool kernels too: the
    parity cases of the small-table kernels must hold there (C1, the small shapes, C5)"""
    import os
    import subprocess
    import sys
    if os.environ.get("PG_TABLES_WIDE"):
        pytest.skip("already inside the wide-kernel run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PG_TABLES_WIDE="1")
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_scan_gpu.py"),
                          "-k", "test_c1_tables or test_small_tables or test_c5_tables_synthetic"], env=env, cwd=root,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-3000:]


@pytest.mark.parametrize("n,A,L,kw", [(17, 4, 3000, {}), (100, 6, 1500, {}), (1000, 4, 300, {}),
                                      (300, 5, 600, dict(min_coverage_depth=0, max_missingness_rate=0.3))])
def test_chisq_many_pools(ctx, n, A, L, kw):
    """tables::chisq beyond 16 pools: one warp per locus (df up to 3,999), unequal pool sizes, pools without coverage"""
    seed = 0xC415 + n
    counts = pb.synth_counts_host(seed, 0, L, n, min(A, 4))
    full = np.zeros((L, A, n), dtype=np.uint32)
    full[:, :min(A, 4)] = counts
    if A > 4:
        rng = np.random.default_rng(seed)
        full[:, 4:] = (rng.random((L, A - 4, n)) < 0.05).astype(np.uint32)
    ps = 3.0 + (np.arange(n) % 4)
    tot = 0.0
    for v in ps:
        tot = tot + v
    fs = _fs(np.array([v / tot for v in ps]), **kw)
    codes = np.arange(A, dtype=np.uint8) if A <= 4 else np.array([0, 1, 2, 3, 4, 5][:A], dtype=np.uint8)
    scan = pb.Scan(ctx, pb.KIND_CHISQ, fs, n, codes)
    dev = scan.run_counts(full)
    scan.close()
    st = H.compare_tables(pb.KIND_CHISQ, full, codes, fs, dev, label=f"chisq n={n}")
    assert st["ok"] > 0.5 * L
    print(st)


@pytest.mark.parametrize("kind", [pb.KIND_CHISQ, pb.KIND_FISHER])
def test_tables_narrow_counts(ctx, kind):
    """u16 / u8 count slabs for the count tests (widened on the device) give the records of the u32 slab, also through
    the streaming submit"""
    n, A, L = 3, 4, 5000
    counts = pb.synth_counts_host(0x7AB1E, 0, L, n, A)
    assert counts.max() < 256
    fs = _fs(np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, kind, fs, n, np.arange(A, dtype=np.uint8))
    ref = scan.run_counts(counts)
    for dt in (np.uint16, np.uint8):
        r = scan.run_counts(counts.astype(dt))
        assert (r.status == ref.status).all() and np.array_equal(r.stats, ref.stats, equal_nan=True)
        assert (r.alleles == ref.alleles).all()
    scan.stream_begin(2048)
    parts, pend = [], []
    for l0 in range(0, L, 2048):
        pend.append(scan.submit_counts(np.ascontiguousarray(counts[l0:l0 + 2048].astype(np.uint8))))
    for t in pend:
        parts.append(scan.collect(t))
    scan.close()
    assert np.array_equal(np.concatenate([p.stats for p in parts]), ref.stats, equal_nan=True)


def test_fixup_path_without_ingest_hints():
    """PG_NOHINT=1 (read once per process by the library) sends every locus whose removed alleles carry reads through
    the fix-up kernel's renormalising path again: the same parity cases must hold there"""
    import os
    import subprocess
    import sys
    if os.environ.get("PG_NOHINT"):
        pytest.skip("already inside the no-hint run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PG_NOHINT="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_scan_gpu.py"), "-m", "gpu", "-q",
                          "-x", "-k", "error_reads or hinted_renormalisation or c1_regression"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=root, env=env, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:]
    assert " passed" in out.stdout


@pytest.mark.parametrize("n,A", [(2, 4), (3, 4), (3, 6), (4, 6), (5, 6)])
def test_fewer_pools_than_coefficients(ctx, n, A):
    """the reference's n < p branch (src/gwas/ols.rs:67-75, 106-111): b = X'(XX')^-1 y, the minimum-norm interpolating
    solution; e'e is rounding noise over a negative n - p, so var is a tiny negative number (t = NaN, p forced to 1) or
    -0 when every residual is exactly zero (p = 0).  beta is checked against the oracle and numpy's pseudo-inverse."""
    rng = np.random.default_rng(1000 * n + A)
    L, k = 600, 2
    counts = rng.integers(1, 60, size=(L, A, n)).astype(np.uint32)
    counts[rng.random(L) < 0.2, A - 1] = 0                      # some loci with one allele fewer
    codes = np.arange(A, dtype=np.uint8)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), remove_ns=False, min_allele_frequency=0.0)
    phen = rng.standard_normal((n, k))
    scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, codes, phen)
    dev = scan.run_counts(counts)
    scan.close()
    ofs = H.oracle_fs(fs)
    orc = pgo.scan_batch(pgo.SCAN_OLS, counts, codes, phen, ofs, 4)
    assert ((orc.status == pgo.FILTERED) == (dev.status == pb.LOCUS_FILTERED)).all()
    assert not (dev.status == pb.LOCUS_UNSUPPORTED).any()
    under = 0
    for l in np.nonzero(orc.status == pgo.OK)[0]:
        m = int(orc.n_out[l])
        if n >= m + 1:
            continue
        under += 1
        assert dev.status[l] == pb.LOCUS_OK and dev.n_out[l] == m and (dev.alleles[l][:m] == orc.allele[l][:m]).all()
        X = H._design(counts[l], codes, ofs)
        cond = np.linalg.cond(X @ X.T)
        bp = np.linalg.pinv(X) @ phen                             # [p, k] minimum-norm solution
        for s in range(m):
            for j in range(k):
                db, ob, xb = dev.stats[l, s, j, 0], orc.stat[l, s, j], bp[1 + s, j]
                scale = max(np.abs(bp[:, j]).max(), 1e-300)
                assert abs(db - xb) <= max(4 * abs(ob - xb), 1e-9 * scale, cond * 1e-15 * scale), (l, s, j, db, ob, xb, cond)
                dp, op = dev.stats[l, s, j, 3], orc.pval[l, s, j]
                assert dp in (0.0, 1.0) and op in (0.0, 1.0)
                dvar = dev.stats[l, s, j, 1]                       # sqrt(var): NaN for var < 0, -0 for var = -0
                assert np.isnan(dvar) or dvar == 0.0
                if dp != op:  # only the exact-zero residual case may differ (rounding noise decides e'e == 0)
                    assert orc.var[l, s, j] == 0.0 or dvar == 0.0, (l, s, j, dp, op)
    assert under > 100
    # the loci with n >= p still take the normal equations
    ge = (orc.status == pgo.OK) & (orc.n_out.astype(int) + 1 < n)
    if ge.any():
        sub = np.nonzero(ge)[0]
        assert np.allclose(dev.stats[sub, 0, 0, 0], orc.stat[sub, 0, 0], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("n,A,k", [(2, 2, 1), (3, 2, 2), (4, 2, 4), (4, 2, 1), (2, 3, 3), (4, 3, 4)])
@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR])
def test_tiny_stage(ctx, kind, n, A, k):
    """two alleles over at most four pools: a ring stage of 256 bytes is smaller than one row of the reduction scratch
    it doubles as (found by tools/fuzz_parity.py: the reduction then never advanced).  Biallelic loci over a handful
    of pools are an everyday input."""
    L = 3000
    rng = np.random.default_rng(n * 10 + A)
    depth = rng.integers(20, 120, (L, n))
    minor = rng.binomial(depth, rng.uniform(0.02, 0.5, (L, 1)))
    counts = np.zeros((L, A, n), dtype=np.uint32)
    counts[:, 0], counts[:, 1] = depth - minor, minor
    if A == 3:
        counts[:, 2] = rng.integers(0, 5, (L, n))
    phen = rng.standard_normal((n, k)) * 3.0 + 10.0
    fs = _fs(np.full(n, 1.0 / n), min_coverage_depth=10, min_allele_frequency=0.01)
    codes = np.array([1, 3, 5][:A], dtype=np.uint8)
    scan = pb.Scan(ctx, kind, fs, n, codes, phen)
    dev = scan.run_counts(counts)
    scan.close()
    st = H.compare_regression(kind, counts, codes, phen, fs, dev, label=f"tiny n={n} A={A} k={k}")
    assert st["kept"] > 0.5 * L
    print(st)
