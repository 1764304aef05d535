"""Pins the CPU oracle (oracle/poolgen_oracle.c) against every known-answer vector the reference's own unit
tests hold for the hot path.  CPU only."""
import numpy as np

from oracle import pgo
from tests import helpers as H

A, T, C, G, N, D = range(6)


def _fs(sizes, maf=0.005):
    return pgo.FilterStats(pool_sizes=np.array(sizes, dtype=np.float64), remove_ns=True, min_coverage_depth=1,
                           min_allele_frequency=maf, max_missingness_rate=0.0)


def test_pearson_known_answer():
    # src/gwas/correlation_test.rs:138-181
    r, p = pgo.pearsons_correlation([0.1, 0.2, 0.3, 0.4, 0.5], [2.0, 1.0, 1.0, 5.0, 2.0])
    assert r == pgo.sensible_round(0.3849001794597505, 7)
    assert p == 0.5223146158470686
    counts = np.array([[1, 9], [2, 8], [3, 7], [4, 6], [5, 5]], dtype=np.uint64)
    res = pgo.correlation(counts, [A, T], np.array([2.0, 1.0, 1.0, 5.0, 2.0]), _fs([20.0] * 5))
    assert pgo.format_corr_lines("Chromosome1", 12345, res) == "Chromosome1,12345,A,0.3,Pheno_0,0.3849,0.5223146158470686\n"
    # NaN handling, correlation_test.rs:182-205
    nan = float("nan")
    r2, _ = pgo.pearsons_correlation([0.1, 0.2, nan, nan, 0.5, 0.6], [0.1, 0.2, nan, nan, 0.5, 0.6])
    assert pgo.sensible_round(r2, 2) == 1.00
    r3, _ = pgo.pearsons_correlation([0.1, 0.2, nan, nan, 0.5, 0.6], [0.1, 0.2, nan, 0.4, nan, 0.6])
    assert pgo.sensible_round(r3, 2) == 1.00
    r4, _ = pgo.pearsons_correlation([nan, nan, nan], [nan, nan, nan])
    assert np.isnan(r4)


def test_chisq_known_answer():
    # src/tables/chisq_test.rs:55-82
    counts = np.array([[0, 20], [20, 0], [0, 20], [20, 0]], dtype=np.uint64)
    res = pgo.chisq(counts, [A, T], _fs([0.2] * 4))
    assert pgo.format_chisq_line("Chromosome1", 12345, res) == "Chromosome1,12345,AT,4,0.7797774084757156\n"


def test_fisher_known_answer():
    # src/tables/fisher_exact_test.rs:137-173
    assert pgo.factorial_log10(5.0) == 2.0791812460476247
    assert pgo.hypergeom_ratio(np.array([[0.0, 3.0], [1.0, 5.0], [2.0, 6.0]]), 19.959563872703743) == 0.24705882352941286
    counts = np.array([[0, 3], [1, 5], [2, 6]], dtype=np.uint64)
    res = pgo.fisher(counts, [T, C], _fs([0.2] * 3))
    assert pgo.format_fisher_line("Chromosome1", 12345, res) == "Chromosome1,12345,TC,0.24705882352941286,0.6073529411764731\n"


def test_sync_parse_filter_sort_known_answer():
    # src/base/sync.rs:1557-1632
    line = "Chromosome1\t456527\tC\t1:0:999:0:4:0\t0:1:2:0:0:0\t0:2:4:0:0:0\t0:1:4:0:0:0\t0:1:6:0:0:0"
    n, chrom, pos, counts = pgo.parse_sync_line(line)
    assert (n, chrom, pos) == (5, "Chromosome1", 456527)
    expect = np.array([[1, 0, 999, 0, 4, 0], [0, 1, 2, 0, 0, 0], [0, 2, 4, 0, 0, 0], [0, 1, 4, 0, 0, 0], [0, 1, 6, 0, 0, 0]], dtype=np.uint64)
    assert (counts == expect).all()
    st, ck, ak = pgo.filter_locus(counts, [A, T, C, G, N, D], _fs([20.0] * 5))
    assert st == pgo.OK and list(ak) == [T, C]
    assert (ck == expect[:, [1, 2]]).all()
    f = pgo.to_frequencies(ck)
    fs_, as_ = pgo.sort_by_allele_freq(f, ak, True)
    assert list(as_) == [C, T]
    # loaded first locus with keep_p_minus_1 (sort, drop the major): T = [0, 1/3, 1/3, 0.2, 1/7]
    assert list(fs_[:, 1]) == [0.0, 0.3333333333333333, 0.3333333333333333, 0.2, 0.14285714285714285]


def test_ols_betas_of_the_commented_vector():
    # src/gwas/ols.rs:534 (stale block: the betas are still valid, the p-values used an older df)
    counts = np.array([[4, 1, 5], [2, 1, 7], [3, 2, 5], [4, 3, 3], [5, 5, 0]], dtype=np.uint64)
    y = np.array([[2.0, 0.5], [1.0, 0.2], [2.0, 0.5], [4.0, 0.0], [5.0, 0.5]])
    res = pgo.ols_iterate(counts, [A, T, D], y, _fs([20.0] * 5))
    assert res.status == pgo.OK and res.alleles == [A, T]
    got = [pgo.round_to_string(v, 6) for v in res.stat.ravel()]
    assert got == ["5.528455", "0.99187", "6.422764", "-0.406504"]
    lines = pgo.format_ols_lines("Chromosome1", 12345, res).split("\n")
    assert lines[0].startswith("Chromosome1,12345,A,0.36,Pheno_0,5.528455,")


def test_rust_display_and_rounding():
    # helpers.rs:103-117 and Rust's f64 Display
    assert pgo.f64_to_string(1e-7) == "0.0000001"
    assert pgo.f64_to_string(4.0) == "4"
    assert pgo.f64_to_string(1e21) == "1000000000000000000000"
    assert pgo.f64_to_string(float("nan")) == "NaN"
    assert pgo.round_to_string(0.36, 8) == "0.36"
    assert pgo.round_to_string(0.3849001794597505, 6) == "0.3849"


def test_c1_keep_counts():
    """tests/test.sync through the restated filter: 6556 of 6674 loci pass with the CLI defaults, 2032 with
    --min-coverage-depth 10 --min-allele-frequency 0.01."""
    c1 = H.load_c1()
    assert c1["counts"].shape == (6674, 6, 5)
    for kw, expect in ((dict(), 6556), (dict(min_coverage_depth=10, min_allele_frequency=0.01), 2032)):
        fs = pgo.FilterStats(pool_sizes=c1["pool_sizes"], **kw)
        r = pgo.scan_batch(pgo.SCAN_OLS, c1["counts"], c1["codes"], c1["phen"], fs, n_threads=4)
        assert int((r.status != pgo.FILTERED).sum()) == expect
        assert not (r.status == pgo.PANIC).any()


def test_statrs_tails_against_scipy():
    from scipy import special, stats
    for df in (3.0, 4.0, 99.0, 998.0):
        for t in (0.01, 0.5, 1.0, 2.5, 6.0):
            p = 2.0 * (1.0 - pgo.students_t_cdf(t, df))
            assert abs(p - 2 * stats.t.sf(t, df)) <= 1e-9 * p + 1e-15
    for a, x in ((3.5, 2.0), (0.5, 0.1), (12.0, 30.0)):
        assert abs(pgo.gamma_lr(a, x) - special.gammainc(a, x)) < 1e-13


def test_tight_cpu_baseline_is_bit_identical_to_the_faithful_one():
    """bench.py reports two CPU baselines: the faithful restatement and the 'tight' one (pool-size total
    hoisted, one inversion per locus, no per-locus allocation).  Same arithmetic, so the same bits -- on the synthetic
    workload, on C1 and with weighted pools."""
    import poolgen_b200 as pb
    from tests import helpers as H
    fields = ("status", "n_out", "allele", "freq_mean", "stat", "var", "t", "pval")
    for n, A, k, L in [(100, 4, 3, 1500), (12, 6, 2, 800), (3, 4, 1, 300)]:
        c = pb.synth_counts_host(0x7167 + n, 0, L, n, A)
        y = pb.synth_phen_host(5, n, k)
        rng = np.random.default_rng(n)
        sizes = rng.integers(5, 60, size=n).astype(float)
        for ps in (np.full(n, 1.0 / n), sizes / sizes.sum()):
            fs = pgo.FilterStats(pool_sizes=ps, min_allele_frequency=0.01)
            a = pgo.scan_batch(pgo.SCAN_OLS, c, np.arange(A, dtype=np.uint8), y, fs, 2)
            b = pgo.scan_batch(pgo.SCAN_OLS, c, np.arange(A, dtype=np.uint8), y, fs, 3, tight=True)
            assert (a.status == pgo.OK).sum() > 0.3 * L
            for f in fields:
                assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), (n, A, f)
    c1 = H.load_c1()
    fs = pgo.FilterStats(pool_sizes=c1["pool_sizes"], min_coverage_depth=10, min_allele_frequency=0.01)
    counts = np.ascontiguousarray(c1["counts"], dtype=np.uint32)
    a = pgo.scan_batch(pgo.SCAN_OLS, counts, c1["codes"], c1["phen"], fs, 2)
    b = pgo.scan_batch(pgo.SCAN_OLS, counts, c1["codes"], c1["phen"], fs, 2, tight=True)
    assert (a.status != pgo.FILTERED).sum() == 2032  # pass the filter; a few then fail the regression
    for f in fields:
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f


def _gwalpha_example():
    """the inputs of the reference's test_gwalpha (src/gwas/gwalpha.rs:399-441)"""
    counts = np.array([5, 2, 6, 2, 2, 7, 3, 2, 5, 4, 3, 3, 5, 5, 0], dtype=np.uint64).reshape(5, 3)
    fmt = np.array([0.2, 0.2, 0.2, 0.2, 0.2, 0.0, 0.1, 0.4, 0.7, 0.9, 0.02, 0.0, 0.9, -np.inf, -np.inf]).reshape(3, 5).T.copy()
    fs = pgo.FilterStats(pool_sizes=np.array([20.0] * 5), remove_ns=True, min_coverage_depth=1,
                         min_allele_frequency=0.005, max_missingness_rate=0.0)
    return counts, np.array([0, 1, 5], dtype=np.uint8), fmt, fs


def test_gwalpha_lines_pin_the_nelder_mead_restatement():
    """src/gwas/gwalpha.rs:392-447 (test_gwalpha): the two output lines of gwalpha_ls and of gwalpha_ml.  They go through
    argmin's Nelder-Mead (1,000-iteration cap), statrs' Beta::cdf and bound_parameters_with_logit, so they pin the
    oracle's restatement of the solver (reflection / expansion / outside and inside contraction / shrink, stable sort,
    standard-deviation stop) to the six printed digits -- for BOTH cost functions and all four searches."""
    counts, alleles, fmt, fs = _gwalpha_example()
    ls = pgo.format_gwalpha_lines("Chromosome1", 12345, pgo.gwalpha(counts, alleles, fmt, fs, "LS"))
    ml = pgo.format_gwalpha_lines("Chromosome1", 12345, pgo.gwalpha(counts, alleles, fmt, fs, "ML"))
    assert ls == "Chromosome1,12345,A,0.353287,Pheno_0,5.816067,Unknown\nChromosome1,12345,T,0.267133,Pheno_0,9.176892,Unknown\n"
    assert ml == "Chromosome1,12345,A,0.353287,Pheno_0,-3.293261,Unknown\nChromosome1,12345,T,0.267133,Pheno_0,-7.098985,Unknown\n"


def test_bound_parameters_with_logit():
    """src/base/helpers.rs:120-129"""
    eps = np.finfo(float).eps
    assert pgo.bound_logit(0.0, eps, 1e9) == eps + (1e9 - eps) / 2.0
    assert pgo.bound_logit(-800.0, eps, 10.0) == eps          # exp(800) = inf: the lower limit
    assert abs(pgo.bound_logit(40.0, eps, 10.0) - 10.0) < 1e-14


def test_mle_iterate_converges_to_the_closed_form():
    """mle_iterate (src/gwas/mle.rs:232-305) minimises (n/2) ln(2 pi s2) + RSS(beta) / s2: at the optimum beta is the OLS
    solution and s2 = 2 RSS / n (the reference's cost has no 1/2 on its second term).  The simplex search is capped at
    1,000 iterations, so the restatement is held to the solver's own convergence, not to 1e-9."""
    import poolgen_b200 as pb
    n, A, k, L = 30, 4, 2, 150
    counts = pb.synth_counts_host(0x31E, 0, L, n, A)
    phen = pb.synth_phen_host(0x31E, n, k)
    fs = pgo.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    codes = np.arange(A, dtype=np.uint8)
    m = pgo.scan_batch(pgo.SCAN_MLE, counts, codes, phen, fs, 4)
    o = pgo.scan_batch(pgo.SCAN_OLS, counts, codes, phen, fs, 4)
    ok = (m.status == pgo.OK) & (o.status == pgo.OK)
    assert ok.sum() > 100 and ((m.status == pgo.FILTERED) == (o.status == pgo.FILTERED)).all()
    assert (m.n_out[ok] == o.n_out[ok]).all() and (m.allele[ok] == o.allele[ok]).all()
    sel = ~np.isnan(o.stat[ok])
    rel = np.abs(m.stat[ok] - o.stat[ok])[sel] / np.maximum(np.abs(o.stat[ok])[sel], np.sqrt(o.var[ok])[sel])
    assert np.median(rel) < 1e-6 and np.quantile(rel, 0.9) < 1e-3
    # v_b = s2 diag((X'X)^-1) with s2 = 2 RSS / n = 2 (n - p) / n times the OLS residual variance
    p_x = m.n_out[ok].astype(float)[:, None, None] + 1.0
    ratio = (m.var[ok] / o.var[ok])[sel]
    expect = np.broadcast_to(2.0 * (n - p_x) / n, m.var[ok].shape)[sel]
    assert np.median(np.abs(ratio / expect - 1.0)) < 1e-5


def test_mle_with_covariate_converges_to_the_closed_form():
    """mle_with_covariate (src/gwas/mle.rs:307-463): with a few covariates the capped simplex search reaches the
    optimum, where the last coefficient is the least-squares one and v_b = (2 RSS / n) [(X'X)^-1]_gg = 2 (n - p) / n
    times the OLS variance.  A constant column has no inverse: NaN, like the reference's Err branch (mle.rs:377-383)."""
    rng = np.random.default_rng(3)
    n, P = 50, 40
    G = np.clip(0.45 + 0.2 * rng.standard_normal((P, n)), 0.0, 1.0)
    G[7] = 0.5
    y = rng.standard_normal((n, 2)) + 2.0 * G[1][:, None]
    for m in (0, 1, 2):
        cov = rng.standard_normal((n, m)) / np.sqrt(n)
        _, b, v, p = pgo.mle_with_covariate(G, y, 0.5, covariates=cov)
        assert np.isnan(b[7]).all() and np.isnan(p[7]).all()
        for c in (0, 1, 5, 30):
            x = np.ones((n, 2 + m))
            x[:, 1:1 + m] = cov
            x[:, 1 + m] = G[c]
            rc, bo, vo, po, _ = pgo.ols(x, y)
            assert rc == 0
            assert np.all(np.abs(b[c] - bo[1 + m]) < 1e-5 * np.sqrt(vo[1 + m]))
            assert np.allclose(v[c] / vo[1 + m], 2.0 * (n - (2 + m)) / n, rtol=1e-5)
            assert np.all((p[c] >= 0) & (p[c] <= 1))


def test_reciprocal_division_is_the_ieee_quotient():
    """The ingest kernel forms the first-stage frequency c / depth with one reciprocal per pool and a correction step
    per allele (pg_ingest.cu); the keep-mask is bit-exact only if that IS the correctly rounded quotient the reference
    computes (src/base/sync.rs:166-192).  The sequence consists of IEEE operations only, so the host reproduces it:
    every c <= d <= 4096 and 2e7 random pairs of 32-bit operands."""
    assert pgo.check_reciprocal_division(4096, 20_000_000) == 0
