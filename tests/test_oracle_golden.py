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
