"""ctypes binding of the CPU oracle (oracle/poolgen_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under poolgen_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpoolgen_oracle.so")
_lib = None

MAX_ALLELES = 6
ALLELE_NAMES = "ATCGND"
FILTERED, OK, FAILED, PANIC = 0, 1, 2, -1
SCAN_OLS, SCAN_CORR, SCAN_CHISQ, SCAN_FISHER = 0, 1, 2, 3
SCAN_MLE, SCAN_GWALPHA_LS, SCAN_GWALPHA_ML = 5, 6, 7


class _FilterStats(C.Structure):
    _fields_ = [
        ("remove_ns", C.c_int),
        ("min_coverage_depth", C.c_uint64),
        ("min_allele_frequency", C.c_double),
        ("max_missingness_rate", C.c_double),
        ("n_pool_sizes", C.c_int),
        ("pool_sizes", C.POINTER(C.c_double)),
    ]


class _LocusResult(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("n_alleles_out", C.c_int),
        ("allele", C.c_uint8 * MAX_ALLELES),
        ("freq_mean", C.c_double * MAX_ALLELES),
        ("stat", C.POINTER(C.c_double)),
        ("var", C.POINTER(C.c_double)),
        ("t", C.POINTER(C.c_double)),
        ("pval", C.POINTER(C.c_double)),
    ]


class _TableResult(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("n_alleles_out", C.c_int),
        ("allele", C.c_uint8 * MAX_ALLELES),
        ("statistic", C.c_double),
        ("pval", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (make -C oracle)."""
    src = os.path.join(_HERE, "poolgen_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        d, dp, i, u8p, u64p = C.c_double, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        for name, res, args in [
            ("pgo_ln_gamma", d, [d]),
            ("pgo_beta_reg", d, [d, d, d]),
            ("pgo_students_t_cdf", d, [d, d]),
            ("pgo_gamma_lr", d, [d, d]),
            ("pgo_chisq_cdf", d, [d, d]),
            ("pgo_lu_inverse", i, [dp, i]),
            ("pgo_lu_det", d, [dp, i]),
            ("pgo_sensible_round", d, [d, i]),
            ("pgo_f64_to_string", i, [d, C.c_char_p, C.c_size_t]),
            ("pgo_round_to_string", i, [d, i, C.c_char_p, C.c_size_t]),
            ("pgo_parse_sync_line", i, [C.c_char_p, C.c_char_p, C.c_size_t, u64p, u64p, i]),
            ("pgo_to_frequencies", None, [u64p, i, i, dp]),
            ("pgo_filter", i, [u64p, u8p, i, C.POINTER(i), C.POINTER(_FilterStats)]),
            ("pgo_sort_by_allele_freq", None, [dp, u8p, i, i, i]),
            ("pgo_ols", i, [dp, i, i, dp, i, dp, dp, dp, dp]),
            ("pgo_ols_iterate", i, [u64p, u8p, i, i, dp, i, C.POINTER(_FilterStats), C.POINTER(_LocusResult)]),
            ("pgo_pearsons_correlation", i, [dp, dp, i, dp, dp]),
            ("pgo_correlation", i, [u64p, u8p, i, i, dp, i, C.POINTER(_FilterStats), C.POINTER(_LocusResult)]),
            ("pgo_chisq", i, [u64p, u8p, i, i, C.POINTER(_FilterStats), C.POINTER(_TableResult)]),
            ("pgo_factorial_log10", d, [d, C.POINTER(i)]),
            ("pgo_hypergeom_ratio", d, [dp, i, d]),
            ("pgo_fisher", i, [u64p, u8p, i, i, C.POINTER(_FilterStats), C.POINTER(_TableResult)]),
            ("pgo_format_ols_lines", i, [C.c_char_p, C.c_uint64, C.POINTER(_LocusResult), i, C.c_char_p, C.c_size_t]),
            ("pgo_format_corr_lines", i, [C.c_char_p, C.c_uint64, C.POINTER(_LocusResult), i, C.c_char_p, C.c_size_t]),
            ("pgo_format_chisq_line", i, [C.c_char_p, C.c_uint64, C.POINTER(_TableResult), C.c_char_p, C.c_size_t]),
            ("pgo_format_fisher_line", i, [C.c_char_p, C.c_uint64, C.POINTER(_TableResult), C.c_char_p, C.c_size_t]),
            ("pgo_bound_logit", d, [d, d, d]),
            ("pgo_check_reciprocal_division", C.c_long, [C.c_uint, C.c_long]),
            ("pgo_mle_iterate", i, [u64p, u8p, i, i, dp, i, C.POINTER(_FilterStats), C.POINTER(_LocusResult)]),
            ("pgo_mle_regress", i, [dp, i, i, dp, dp, dp, dp]),
            ("pgo_gwalpha", i, [u64p, u8p, i, i, dp, i, i, C.POINTER(_FilterStats), C.POINTER(_LocusResult)]),
            ("pgo_format_mle_lines", i, [C.c_char_p, C.c_uint64, C.POINTER(_LocusResult), i, C.c_char_p, C.c_size_t]),
            ("pgo_format_gwalpha_lines", i, [C.c_char_p, C.c_uint64, C.POINTER(_LocusResult), C.c_char_p, C.c_size_t]),
            ("pgo_scan_batch", i, [i, C.POINTER(C.c_uint32), C.c_int64, i, i, u8p, dp, i,
                                   C.POINTER(_FilterStats), i, C.POINTER(C.c_int8), u8p, u8p, dp, dp, dp, dp, dp]),
            ("pgo_scan_batch_tight", i, [i, C.POINTER(C.c_uint32), C.c_int64, i, i, u8p, dp, i,
                                         C.POINTER(_FilterStats), i, C.POINTER(C.c_int8), u8p, u8p, dp, dp, dp, dp, dp]),
        ]:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


@dataclass
class FilterStats:
    """FilterStats of the reference (structs_and_traits.rs:69-78), sync-path fields only."""
    pool_sizes: np.ndarray
    remove_ns: bool = True
    min_coverage_depth: int = 1
    min_allele_frequency: float = 0.001
    max_missingness_rate: float = 0.0
    _keep: object = field(default=None, repr=False)

    def c(self) -> _FilterStats:
        ps = np.ascontiguousarray(self.pool_sizes, dtype=np.float64)
        self._keep = ps
        return _FilterStats(int(self.remove_ns), int(self.min_coverage_depth),
                            float(self.min_allele_frequency), float(self.max_missingness_rate),
                            int(ps.size), _dp(ps))


# ---- scalar numerics ------------------------------------------------------------------
def ln_gamma(x): return lib().pgo_ln_gamma(float(x))
def beta_reg(a, b, x): return lib().pgo_beta_reg(float(a), float(b), float(x))
def students_t_cdf(x, df): return lib().pgo_students_t_cdf(float(x), float(df))
def gamma_lr(a, x): return lib().pgo_gamma_lr(float(a), float(x))
def chisq_cdf(x, df): return lib().pgo_chisq_cdf(float(x), float(df))
def sensible_round(x, d): return lib().pgo_sensible_round(float(x), int(d))
def factorial_log10(x): return lib().pgo_factorial_log10(float(x), None)


def hypergeom_ratio(counts, lp):
    c = np.ascontiguousarray(counts, dtype=np.float64).ravel()
    return lib().pgo_hypergeom_ratio(_dp(c), int(c.size), float(lp))


def lu_inverse(a):
    """inverse of a square matrix through the dgetrf/dgetri restatement (C-layout in, as ndarray-linalg does)."""
    m = np.array(a, dtype=np.float64, order="C")
    info = lib().pgo_lu_inverse(_dp(m), m.shape[0])
    return m, info


def f64_to_string(x):
    buf = C.create_string_buffer(400)
    lib().pgo_f64_to_string(float(x), buf, 400)
    return buf.value.decode()


def round_to_string(x, d):
    buf = C.create_string_buffer(400)
    lib().pgo_round_to_string(float(x), int(d), buf, 400)
    return buf.value.decode()


# ---- sync.rs --------------------------------------------------------------------------
def parse_sync_line(line: str, max_pools: int = 4096):
    counts = np.zeros((max_pools, 6), dtype=np.uint64)
    chrom = C.create_string_buffer(256)
    pos = C.c_uint64(0)
    n = lib().pgo_parse_sync_line(line.encode(), chrom, 256, C.byref(pos), _u64p(counts), max_pools)
    if n <= 0:
        return n, None, None, None
    return n, chrom.value.decode(), int(pos.value), counts[:n].copy()


def to_frequencies(counts):
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    f = np.empty(c.shape, dtype=np.float64)
    lib().pgo_to_frequencies(_u64p(c), c.shape[0], c.shape[1], _dp(f))
    return f


def filter_locus(counts, alleles, fs: FilterStats):
    """LocusCounts::filter.  Returns (status, counts_kept (n x p'), alleles_kept)."""
    c = np.array(counts, dtype=np.uint64, order="C")
    a = np.array(alleles, dtype=np.uint8)
    n, p0 = c.shape
    p = C.c_int(p0)
    fsc = fs.c()
    st = lib().pgo_filter(_u64p(c), _u8p(a), n, C.byref(p), C.byref(fsc))
    pk = p.value
    return st, c.ravel()[: n * pk].reshape(n, pk).copy(), a[:pk].copy()


def sort_by_allele_freq(freq, alleles, decreasing=True):
    f = np.array(freq, dtype=np.float64, order="C")
    a = np.array(alleles, dtype=np.uint8)
    lib().pgo_sort_by_allele_freq(_dp(f), _u8p(a), f.shape[0], f.shape[1], int(decreasing))
    return f, a


# ---- gwas -----------------------------------------------------------------------------
def ols(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None].copy()
    n, p = x.shape
    k = y.shape[1]
    beta, var, pval, t = (np.full((p, k), np.nan) for _ in range(4))
    rc = lib().pgo_ols(_dp(x), n, p, _dp(y), k, _dp(beta), _dp(var), _dp(pval), _dp(t))
    return rc, beta, var, pval, t


@dataclass
class LocusResult:
    status: int
    alleles: list
    freq_mean: np.ndarray
    stat: np.ndarray  # (n_alleles_out, k)
    var: np.ndarray
    t: np.ndarray
    pval: np.ndarray
    _raw: object = field(default=None, repr=False)


def _locus_call(fn, counts, alleles, phen, fs: FilterStats):
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    a = np.ascontiguousarray(alleles, dtype=np.uint8)
    y = np.ascontiguousarray(phen, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None].copy()
    n, p = c.shape
    k = y.shape[1]
    bufs = [np.full(MAX_ALLELES * k, np.nan) for _ in range(4)]
    r = _LocusResult()
    r.stat, r.var, r.t, r.pval = (_dp(b) for b in bufs)
    fsc = fs.c()
    st = fn(_u64p(c), _u8p(a), n, p, _dp(y), k, C.byref(fsc), C.byref(r))
    m = r.n_alleles_out
    out = LocusResult(st, [int(r.allele[i]) for i in range(m)],
                      np.array([r.freq_mean[i] for i in range(m)]),
                      *(b[: m * k].reshape(m, k).copy() for b in bufs))
    out._raw = (r, bufs, k)
    return out


def ols_iterate(counts, alleles, phen, fs): return _locus_call(lib().pgo_ols_iterate, counts, alleles, phen, fs)
def mle_iterate(counts, alleles, phen, fs): return _locus_call(lib().pgo_mle_iterate, counts, alleles, phen, fs)


def mle_regress(x, y):
    """mle(x, y, false) for one phenotype (gwas/mle.rs:193-230): (rc, beta [p], var [p], pval [p])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    n, p = x.shape
    beta, var, pval = (np.full(p, np.nan) for _ in range(3))
    rc = lib().pgo_mle_regress(_dp(x), n, p, _dp(y), _dp(beta), _dp(var), _dp(pval))
    return rc, beta, var, pval


def check_reciprocal_division(d_max, n_random):
    """mismatches between RN(q + (c - q d) r), r = RN(1 / d), q = RN(c r) and the IEEE quotient c / d"""
    return int(lib().pgo_check_reciprocal_division(int(d_max), int(n_random)))


def bound_logit(x, lower, upper): return lib().pgo_bound_logit(float(x), float(lower), float(upper))


def gwalpha(counts, alleles, phen_fmt, fs, method="LS"):
    """gwalpha_ls / gwalpha_ml (gwas/gwalpha.rs:282-386); phen_fmt: the gwalpha_fmt matrix [rows, 3] (column 0 bins,
    column 1 q, column 2 = sig, min, max, then -inf).  stat[:, 0] = alpha per allele."""
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    a = np.ascontiguousarray(alleles, dtype=np.uint8)
    y = np.ascontiguousarray(phen_fmt, dtype=np.float64)
    assert y.ndim == 2 and y.shape[1] == 3
    n, p = c.shape
    bufs = [np.full(MAX_ALLELES, np.nan) for _ in range(4)]
    r = _LocusResult()
    r.stat, r.var, r.t, r.pval = (_dp(b) for b in bufs)
    fsc = fs.c()
    st = lib().pgo_gwalpha(_u64p(c), _u8p(a), n, p, _dp(y), int(y.shape[0]), 0 if method == "LS" else 1, C.byref(fsc), C.byref(r))
    m = r.n_alleles_out
    out = LocusResult(st, [int(r.allele[i]) for i in range(m)], np.array([r.freq_mean[i] for i in range(m)]),
                      *(b[:m].reshape(m, 1).copy() for b in bufs))
    out._raw = (r, bufs, 1)
    return out


def format_mle_lines(chrom, pos, res: LocusResult):
    r, _, k = res._raw
    buf = C.create_string_buffer(1 << 14)
    lib().pgo_format_mle_lines(chrom.encode(), int(pos), C.byref(r), k, buf, 1 << 14)
    return buf.value.decode()


def format_gwalpha_lines(chrom, pos, res: LocusResult):
    r, _, _ = res._raw
    buf = C.create_string_buffer(1 << 14)
    lib().pgo_format_gwalpha_lines(chrom.encode(), int(pos), C.byref(r), buf, 1 << 14)
    return buf.value.decode()
def correlation(counts, alleles, phen, fs): return _locus_call(lib().pgo_correlation, counts, alleles, phen, fs)


def pearsons_correlation(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    r, p = C.c_double(), C.c_double()
    lib().pgo_pearsons_correlation(_dp(x), _dp(y), int(x.size), C.byref(r), C.byref(p))
    return r.value, p.value


def format_ols_lines(chrom, pos, res: LocusResult):
    r, _, k = res._raw
    buf = C.create_string_buffer(1 << 14)
    lib().pgo_format_ols_lines(chrom.encode(), int(pos), C.byref(r), k, buf, 1 << 14)
    return buf.value.decode()


def format_corr_lines(chrom, pos, res: LocusResult):
    r, _, k = res._raw
    buf = C.create_string_buffer(1 << 14)
    lib().pgo_format_corr_lines(chrom.encode(), int(pos), C.byref(r), k, buf, 1 << 14)
    return buf.value.decode()


@dataclass
class TableResult:
    status: int
    alleles: list
    statistic: float
    pval: float
    _raw: object = field(default=None, repr=False)


def _table_call(fn, counts, alleles, fs):
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    a = np.ascontiguousarray(alleles, dtype=np.uint8)
    r = _TableResult()
    fsc = fs.c()
    st = fn(_u64p(c), _u8p(a), c.shape[0], c.shape[1], C.byref(fsc), C.byref(r))
    return TableResult(st, [int(r.allele[i]) for i in range(r.n_alleles_out)], r.statistic, r.pval, r)


def chisq(counts, alleles, fs): return _table_call(lib().pgo_chisq, counts, alleles, fs)
def fisher(counts, alleles, fs): return _table_call(lib().pgo_fisher, counts, alleles, fs)


def format_chisq_line(chrom, pos, res: TableResult):
    buf = C.create_string_buffer(1024)
    lib().pgo_format_chisq_line(chrom.encode(), int(pos), C.byref(res._raw), buf, 1024)
    return buf.value.decode()


def format_fisher_line(chrom, pos, res: TableResult):
    buf = C.create_string_buffer(1024)
    lib().pgo_format_fisher_line(chrom.encode(), int(pos), C.byref(res._raw), buf, 1024)
    return buf.value.decode()


# ---- batch ----------------------------------------------------------------------------
@dataclass
class BatchResult:
    status: np.ndarray      # int8 [L]
    n_out: np.ndarray       # uint8 [L]
    allele: np.ndarray      # uint8 [L, 6]  (0xff = unused)
    freq_mean: np.ndarray   # f64 [L, 6]
    stat: np.ndarray        # f64 [L, 6, k]
    var: np.ndarray
    t: np.ndarray
    pval: np.ndarray


def scan_batch(kind, counts_packed, allele_codes, phen, fs: FilterStats, n_threads=1, tight=False) -> BatchResult:
    """counts_packed: uint32 [L, A, n] (allele-major, pools contiguous).  tight=True: the allocation-free ols_iter
    variant with the pool-size total hoisted and one inversion per locus (bit-identical records)."""
    cp = np.ascontiguousarray(counts_packed, dtype=np.uint32)
    L, A, n = cp.shape
    codes = np.ascontiguousarray(allele_codes, dtype=np.uint8)
    assert codes.size == A
    if phen is None:
        y = np.zeros((n, 1))
    else:
        y = np.ascontiguousarray(phen, dtype=np.float64)
        if y.ndim == 1:
            y = y[:, None].copy()
    k = y.shape[1]
    k_arg = k
    if kind in (SCAN_GWALPHA_LS, SCAN_GWALPHA_ML):  # phen = gwalpha_fmt matrix [rows, 3]; one alpha per allele
        assert y.shape[1] == 3
        k_arg, k = int(y.shape[0]), 1
    status = np.zeros(L, dtype=np.int8)
    n_out = np.zeros(L, dtype=np.uint8)
    allele = np.full((L, MAX_ALLELES), 0xFF, dtype=np.uint8)
    fm = np.full((L, MAX_ALLELES), np.nan)
    stat, var, t, pval = (np.full((L, MAX_ALLELES, k), np.nan) for _ in range(4))
    fsc = fs.c()
    fn = lib().pgo_scan_batch_tight if tight else lib().pgo_scan_batch
    rc = fn(int(kind), cp.ctypes.data_as(C.POINTER(C.c_uint32)), L, n, A, _u8p(codes),
                              _dp(y), k_arg, C.byref(fsc), int(n_threads),
                              status.ctypes.data_as(C.POINTER(C.c_int8)), _u8p(n_out), _u8p(allele),
                              _dp(fm), _dp(stat), _dp(var), _dp(t), _dp(pval))
    assert rc == 0
    return BatchResult(status, n_out, allele, fm, stat, var, t, pval)


# ---- ols_iter_with_kinship: LoadAll (sync.rs:973-1179) + ols_with_covariate (gwas/ols.rs:278-436) -------------------
def load_columns(counts, alleles, fs: FilterStats, keep_p_minus_1: bool = False):
    """LoadAll::per_chunk_load + into_genotypes_and_phenotypes for a list of loci (counts [L, n, A] u64):
    returns (G [P, n] allele columns = intercept_and_allele_frequencies[:, 1..] transposed, labels [(locus, allele)])."""
    cols, labels = [], []
    for l in range(len(counts)):
        st, ck, ak = filter_locus(counts[l], alleles, fs)
        if st != OK:
            continue
        f = to_frequencies(ck)
        if keep_p_minus_1:
            f, ak = sort_by_allele_freq(f, ak, True)
            f, ak = f[:, 1:], ak[1:]
        for j in range(f.shape[1]):
            cols.append(f[:, j].copy())
            labels.append((l, int(ak[j])))
    n = counts.shape[1]
    return (np.array(cols) if cols else np.zeros((0, n))), labels


def select_eigenvectors(eigvals_desc, threshold):
    """the PC-count rule of gwas/ols.rs:297-311 on eigenvalues sorted from high to low (the reference's assumption)"""
    n = len(eigvals_desc)
    s = 0.0
    for v in eigvals_desc:
        s = s + v
    cum = [v / s for v in eigvals_desc]
    m = n
    for i in range(1, n):
        cum[i] = cum[i - 1] + cum[i]
        if cum[i - 1] >= threshold and (i - 1) < m:
            m = i - 1
    return m


def ols_with_covariate(G_cols, phen, threshold, columns=None, return_eig=False):
    """G_cols [P, n] allele columns, phen [n, k].  Returns (m, beta [P, k], var [P, k], pval [P, k]) where each entry is
    the LAST coefficient of ols([1 | PCs | g], y) (gwas/ols.rs:340-370); NaN where the regression fails.
    The eigen-decomposition uses numpy's symmetric solver with eigenvalues sorted from high to low; the reference calls
    MKL dgeev and ASSUMES that order (gwas/ols.rs:296) -- parity unpinned for the order, pinned for everything else by
    the invariance of the last coefficient to the basis of span(PCs).
    columns: optional subset of column ordinals to regress (the records of the others stay NaN) -- the kinship matrix is
    always formed from ALL columns.  return_eig: additionally return (eigenvalues high to low, eigenvectors)."""
    G = np.ascontiguousarray(G_cols, dtype=np.float64)
    P, n = G.shape
    y = np.ascontiguousarray(phen, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None].copy()
    k = y.shape[1]
    K = (G.T @ G) / float(P)
    w, V = np.linalg.eigh(K)
    w, V = w[::-1], V[:, ::-1]
    m = select_eigenvectors(list(w), threshold)
    cov = V[:, :m]
    beta, var, pval = (np.full((P, k), np.nan) for _ in range(3))
    for c in (range(P) if columns is None else columns):
        x = np.ones((n, 2 + m))
        x[:, 1:1 + m] = cov
        x[:, 1 + m] = G[c]
        rc, b, v, p, _ = ols(x, y)
        if rc == 0:
            beta[c], var[c], pval[c] = b[1 + m], v[1 + m], p[1 + m]
    if return_eig:
        return m, beta, var, pval, w, V
    return m, beta, var, pval


def mle_with_covariate(G_cols, phen, threshold, columns=None, covariates=None):
    """mle_with_covariate (gwas/mle.rs:307-463): like ols_with_covariate above with mle() in place of ols() -- each entry
    the last coefficient of the maximum-likelihood fit of y on [1 | PCs | g] (Nelder-Mead on sigma2 and the betas), its
    v_b and p (t = beta / v_b, df n - 1, mle.rs:160-185); NaN where the regression fails.  covariates: use these columns
    [n, m] instead of the eigenvectors (the simplex search is NOT invariant to the basis of span(PCs), so a parity test
    hands both sides the same columns).  Returns (m, beta [P, k], var, pval)."""
    G = np.ascontiguousarray(G_cols, dtype=np.float64)
    P, n = G.shape
    y = np.ascontiguousarray(phen, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None].copy()
    k = y.shape[1]
    if covariates is None:
        K = (G.T @ G) / float(P)
        w, V = np.linalg.eigh(K)
        w, V = w[::-1], V[:, ::-1]
        m = select_eigenvectors(list(w), threshold)
        cov = V[:, :m]
    else:
        cov = np.ascontiguousarray(covariates, dtype=np.float64).reshape(n, -1)
        m = cov.shape[1]
    beta, var, pval = (np.full((P, k), np.nan) for _ in range(3))
    for c in (range(P) if columns is None else columns):
        x = np.ones((n, 2 + m))
        x[:, 1:1 + m] = cov
        x[:, 1 + m] = G[c]
        for j in range(k):
            rc, b, v, p = mle_regress(x, y[:, j])
            if rc == 0:
                beta[c, j], var[c, j], pval[c, j] = b[1 + m], v[1 + m], p[1 + m]
    return m, beta, var, pval
