"""The CSV writer of the C ABI (pg_format_*: host code) against the oracle's restatement of Rust's
f64 Display + src/base/helpers.rs:103-117 and against the output lines the reference's own tests pin.  CPU only: the
writer formats records, it computes nothing on them."""
import struct

import numpy as np
import pytest

import poolgen_b200 as pb
from oracle import pgo
from tests import helpers as H

A, T, C, G, N, D = range(6)


def _fs(sizes, maf=0.005):
    return pgo.FilterStats(pool_sizes=np.array(sizes, dtype=np.float64), remove_ns=True, min_coverage_depth=1,
                           min_allele_frequency=maf, max_missingness_rate=0.0)


def _records_of_locus(res, k):
    """oracle LocusResult / TableResult -> ScanResults of one locus"""
    if isinstance(res, pgo.TableResult):
        m = len(res.alleles)
        al = np.full((1, 6), 0xFF, np.uint8)
        al[0, :m] = res.alleles
        st = np.full((1, 1, 1, 4), np.nan)
        st[0, 0, 0, 0], st[0, 0, 0, 3] = res.statistic, res.pval
        return pb.ScanResults(np.array([res.status], np.uint8), np.array([m], np.uint8), al, np.full((1, 1), np.nan), st)
    m = len(res.alleles)
    al = np.full((1, 6), 0xFF, np.uint8)
    al[0, :m] = res.alleles
    fm = np.full((1, 5), np.nan)
    fm[0, :m] = res.freq_mean
    st = np.full((1, 5, k, 4), np.nan)
    st[0, :m, :, 0] = res.stat
    st[0, :m, :, 3] = res.pval
    return pb.ScanResults(np.array([res.status], np.uint8), np.array([m], np.uint8), al, fm, st)


def _rows(kind, rec, chrom="Chromosome1", pos=12345):
    return pb.format_rows(kind, rec, [pos], chr_names=[chrom], chr_index=[0], n_threads=1).decode()


def test_pinned_reference_lines():
    # src/gwas/correlation_test.rs:138-181
    counts = np.array([[1, 9], [2, 8], [3, 7], [4, 6], [5, 5]], dtype=np.uint64)
    res = pgo.correlation(counts, [A, T], np.array([2.0, 1.0, 1.0, 5.0, 2.0]), _fs([20.0] * 5))
    assert _rows(pb.KIND_CORR, _records_of_locus(res, 1)) == "Chromosome1,12345,A,0.3,Pheno_0,0.3849,0.5223146158470686\n"
    # src/tables/chisq_test.rs:55-82
    counts = np.array([[0, 20], [20, 0], [0, 20], [20, 0]], dtype=np.uint64)
    res = pgo.chisq(counts, [A, T], _fs([0.2] * 4))
    assert _rows(pb.KIND_CHISQ, _records_of_locus(res, 1)) == "Chromosome1,12345,AT,4,0.7797774084757156\n"
    # src/tables/fisher_exact_test.rs:137-173
    counts = np.array([[0, 3], [1, 5], [2, 6]], dtype=np.uint64)
    res = pgo.fisher(counts, [T, C], _fs([0.2] * 3))
    assert _rows(pb.KIND_FISHER, _records_of_locus(res, 1)) == "Chromosome1,12345,TC,0.24705882352941286,0.6073529411764731\n"
    # src/gwas/ols.rs:534 (the betas of the stale block)
    counts = np.array([[4, 1, 5], [2, 1, 7], [3, 2, 5], [4, 3, 3], [5, 5, 0]], dtype=np.uint64)
    y = np.array([[2.0, 0.5], [1.0, 0.2], [2.0, 0.5], [4.0, 0.0], [5.0, 0.5]])
    res = pgo.ols_iterate(counts, [A, T, D], y, _fs([20.0] * 5))
    text = _rows(pb.KIND_OLS, _records_of_locus(res, 2))
    assert text == pgo.format_ols_lines("Chromosome1", 12345, res)
    assert text.split("\n")[0].startswith("Chromosome1,12345,A,0.36,Pheno_0,5.528455,")
    assert [ln.split(",")[5] for ln in text.strip().split("\n")] == ["5.528455", "0.99187", "6.422764", "-0.406504"]


def test_headers():
    assert pb.format_header(pb.KIND_OLS) == b"#chr,pos,alleles,freq,phenotype,statistic,pvalue\n"    # sync.rs:950
    assert pb.format_header(pb.KIND_CORR) == b"#chr,pos,alleles,freq,phenotype,statistic,pvalue\n"
    assert pb.format_header(pb.KIND_CHISQ) == b"#chr,pos,alleles,statistic,pvalue\n"                 # sync.rs:766
    assert pb.format_header(pb.KIND_FISHER) == b"#chr,pos,alleles,statistic,pvalue\n"
    assert pb.format_header(pb.capi.KIND_OLS_KINSHIP) == b"#chr,pos,alleles,phenotype,statistic,pvalue\n"  # ols.rs:409


def test_numbers_against_the_oracle():
    """f64::to_string and parse_f64_roundup_and_own over magnitudes, signs, ties, specials and random bit patterns"""
    rng = np.random.default_rng(7)
    xs = [0.0, -0.0, 1.0, -1.0, 4.0, 0.36, 0.3, 1e-7, 1e21, 1e22, 1e23, 5e-324, 1.7976931348623157e308, 0.1 + 0.2,
          0.0000005, 0.0000015, 2.5e-7, -2.5e-7, -1e-9, 1e15, 1e16, 123456789.5, 0.9999995, 0.99999949999, 12345.678905,
          float("nan"), float("inf"), float("-inf"), 2.220446049250313e-16, 1.0 - 2 ** -53, 9.5e14, 8.9e8, 9.1e8]
    xs += list(10.0 ** rng.uniform(-14, 10, 4000) * rng.choice([-1.0, 1.0], 4000))
    xs += list(rng.uniform(0, 1, 4000))
    xs += list(np.round(rng.uniform(-50, 50, 2000), 3))
    xs += [struct.unpack("<d", struct.pack("<Q", int(b)))[0] for b in rng.integers(0, 2 ** 63, 3000, dtype=np.uint64) * 2]
    for x in xs:
        assert pb.format_f64(x, 0) == pgo.f64_to_string(x), x
        for d in (6, 7, 8, 12):
            assert pb.format_f64(x, d) == pgo.round_to_string(x, d), (x, d)


@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR, pb.KIND_CHISQ, pb.KIND_FISHER])
def test_c1_rows_match_the_oracle_lines(kind):
    """C1 (tests/test.sync + tests/test.csv): oracle records -> product writer (threads over locus ranges) equals the
    oracle's line formatter locus by locus, chromosome names from a name table"""
    c1 = H.load_c1()
    L = 1500
    counts = c1["counts"][:L]
    regression = kind in (pb.KIND_OLS, pb.KIND_CORR)
    if regression:
        fs = pgo.FilterStats(pool_sizes=c1["pool_sizes"])
        phen = c1["phen"]
        k = phen.shape[1]
    else:
        fs = pgo.FilterStats(pool_sizes=c1["pool_sizes"], min_coverage_depth=1)
        phen, k = None, 1
    okind = {pb.KIND_OLS: pgo.SCAN_OLS, pb.KIND_CORR: pgo.SCAN_CORR, pb.KIND_CHISQ: pgo.SCAN_CHISQ,
             pb.KIND_FISHER: pgo.SCAN_FISHER}[kind]
    r = pgo.scan_batch(okind, counts, c1["codes"], phen, fs, n_threads=4)
    S = 5 if regression else 1
    st = np.full((L, S, k, 4), np.nan)
    st[..., 0] = r.stat[:, :S, :k]
    st[..., 3] = r.pval[:, :S, :k]
    status = np.where(r.status < 0, pb.LOCUS_PANIC, r.status).astype(np.uint8)
    rec = pb.ScanResults(status, r.n_out, r.allele, np.ascontiguousarray(r.freq_mean[:, :S]), st)
    names = [str(s) for s in c1["chrom_names"]]
    got = pb.format_rows(kind, rec, c1["pos"][:L], chr_names=names, chr_index=c1["chrom_idx"][:L], n_threads=3).decode()
    expect = []
    for l in range(L):
        c = counts[l].T.astype(np.uint64)
        chrom, pos = names[int(c1["chrom_idx"][l])], int(c1["pos"][l])
        if kind == pb.KIND_OLS:
            expect.append(pgo.format_ols_lines(chrom, pos, pgo.ols_iterate(c, c1["codes"], phen, fs)))
        elif kind == pb.KIND_CORR:
            expect.append(pgo.format_corr_lines(chrom, pos, pgo.correlation(c, c1["codes"], phen, fs)))
        elif kind == pb.KIND_CHISQ:
            expect.append(pgo.format_chisq_line(chrom, pos, pgo.chisq(c, c1["codes"], fs)))
        else:
            expect.append(pgo.format_fisher_line(chrom, pos, pgo.fisher(c, c1["codes"], fs)))
    expect = "".join(expect)
    assert got.count("\n") == expect.count("\n") > 1000
    assert got == expect


def test_rows_from_sync_text_labels_and_capacity():
    """chromosome names cut out of the sync text by line offset; too small a buffer reports the bytes needed"""
    import ctypes as Ct
    text = b"#comment\nchrA\t10\tN\t1:2:0:0:0:0\t3:1:0:0:0:0\nscaffold_7\t123456789\tN\t5:5:0:0:0:0\t1:9:0:0:0:0\n"
    off = [9, text.index(b"scaffold_7")]
    al = np.full((2, 6), 0xFF, np.uint8)
    al[:, 0] = [A, T]
    st = np.full((2, 5, 1, 4), np.nan)
    st[:, 0, 0, 0] = [1.23456789, -0.5]
    st[:, 0, 0, 3] = [0.05, 1.0]
    rec = pb.ScanResults(np.array([1, 1], np.uint8), np.array([1, 1], np.uint8), al, np.array([[0.25] + [np.nan] * 4] * 2), st)
    got = pb.format_rows(pb.KIND_OLS, rec, [10, 123456789], text=text, line_offsets=off).decode()
    assert got == "chrA,10,A,0.25,Pheno_0,1.234568,0.05\nscaffold_7,123456789,T,0.25,Pheno_0,-0.5,1\n"
    rc_, keep = rec.to_c()
    lab = pb.capi._RowLabels()
    pos = np.array([10, 123456789], np.uint64)
    offs = np.array(off, np.uint64)
    lab.positions = pos.ctypes.data_as(Ct.POINTER(Ct.c_uint64))
    lab.text = text
    lab.line_offsets = offs.ctypes.data_as(Ct.POINTER(Ct.c_uint64))
    need = Ct.c_size_t()
    buf = Ct.create_string_buffer(8)
    rc = pb.capi.lib().pg_format_rows(pb.KIND_OLS, Ct.byref(rc_), Ct.byref(lab), 1, buf, 8, Ct.byref(need))
    assert rc != 0 and need.value == len(got)


def test_kinship_rows():
    """src/gwas/ols.rs:410-433: phenotype outer, column inner, full precision, NaN literal, labels by column ordinal"""
    beta = np.array([[0.5, np.nan, -1e-7], [3.0, 0.1 + 0.2, 1e21]])
    pval = np.array([[0.01, np.nan, 1.0], [2.220446049250313e-16, 0.5, 0.25]])
    chrom = ["intercept", "chr1", "chr1"]
    pos = [0, 100, 100]
    allele = ["intercept", "A", "T"]
    got = pb.format_kinship_rows(chrom, pos, allele, beta, pval, n_threads=2).decode().split("\n")
    assert got[0] == "intercept,0,intercept,Pheno_0,0.5,0.01"
    assert got[1] == "chr1,100,A,Pheno_0,NaN,NaN"
    assert got[2] == "chr1,100,T,Pheno_0,-0.0000001,1"
    assert got[3] == "intercept,0,intercept,Pheno_1,3,0.0000000000000002220446049250313"
    assert got[4] == "chr1,100,A,Pheno_1,0.30000000000000004,0.5"
    assert got[5] == "chr1,100,T,Pheno_1,1000000000000000000000,0.25"
    assert got[6] == "" and len(got) == 7


def _oracle_sync2csv_rows(cols, labels, chroms, pos):
    """rows of SaveCsv::write_csv (src/base/sync.rs:1243-1260) from the oracle's loader output, loci in the order of
    LoadAll::load (stable sort by chromosome string, then position, src/base/sync.rs:1092-1101)"""
    loci = sorted(range(len(chroms)), key=lambda l: (chroms[l].encode(), pos[l]))   # Python's sort is stable
    by_locus = {}
    for c, (l, a) in enumerate(labels):
        by_locus.setdefault(l, []).append((c, a))
    out = []
    for l in loci:
        for c, a in by_locus.get(l, []):
            out.append(",".join([chroms[l], str(pos[l]), "ATCGND"[a]] + [pgo.round_to_string(v, 6) for v in cols[c]]) + "\n")
    return "".join(out)


@pytest.mark.parametrize("keep_p_minus_1", [False, True])
def test_sync2csv_rows_c1(keep_p_minus_1):
    """sync2csv over C1: the oracle's LoadAll columns through pg_sort_loci + pg_format_frequency_rows; the chromosome
    order is scrambled first so that the sort does something"""
    c1 = H.load_c1()
    L = 1200
    counts = c1["counts"][:L]
    fs = pgo.FilterStats(pool_sizes=c1["pool_sizes"])
    cols, labels = pgo.load_columns(counts.transpose(0, 2, 1).astype(np.uint64), c1["codes"], fs, keep_p_minus_1)
    names = ["chr10", "chr2", "chr1", "Chr1", "chr1_b"]
    idx = (np.arange(L) * 7) % len(names)
    pos = (np.arange(L) * 7919) % 5000 + 1
    chroms = [names[i] for i in idx]
    order = pb.sort_loci(pos, chr_names=names, chr_index=idx)
    assert sorted(order) == list(range(L))
    got = pb.format_frequency_rows(cols, [l for l, _ in labels], [a for _, a in labels], pos, locus_order=order,
                                   n_threads=3, chr_names=names, chr_index=idx).decode()
    expect = _oracle_sync2csv_rows(cols, labels, chroms, [int(p) for p in pos])
    assert got == expect and got.count("\n") == len(labels) > 1000
    # as stored (no order): the loader's own sequence
    plain = pb.format_frequency_rows(cols[:5], [l for l, _ in labels[:5]], [a for _, a in labels[:5]], pos,
                                     chr_names=names, chr_index=idx).decode().split("\n")
    assert plain[0].startswith(f"{chroms[labels[0][0]]},{pos[labels[0][0]]},{'ATCGND'[labels[0][1]]},")
    assert pb.format_frequency_header(["Pop1", "Pop2", "Pop3"]) == b"#chr,pos,allele,Pop1,Pop2,Pop3\n"


@pytest.mark.parametrize("n,L", [(100, 1200), (1000, 150), (5, 1500)])
def test_exact_p_option_reprints_the_reference_digits(n, L):
    """PG_FORMAT_EXACT_P: the writer re-derives p from the record's t statistic with the reference's own arithmetic
    (statrs' continued fraction, host libm).  Records whose p was deliberately spoiled in the 10th significant digit
    (the device's table is smooth to ~1e-12) come out with the oracle's lines, text for text."""
    k = 2
    counts = pb.synth_counts_host(0xE1AC7 + n, 0, L, n, 4)
    phen = pb.synth_phen_host(0xE1AC7, n, k)
    fs = pgo.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    codes = np.arange(4, dtype=np.uint8)
    r = pgo.scan_batch(pgo.SCAN_OLS, counts, codes, phen, fs, n_threads=4)
    S = 3
    st = np.full((L, S, k, 4), np.nan)
    st[..., 0] = r.stat[:, :S, :k]
    st[..., 2] = r.t[:, :S, :k]
    st[..., 3] = r.pval[:, :S, :k] * (1.0 + 3e-10)
    status = np.where(r.status < 0, pb.LOCUS_PANIC, r.status).astype(np.uint8)
    rec = pb.ScanResults(status, r.n_out, r.allele, np.ascontiguousarray(r.freq_mean[:, :S]), st)
    pos = np.arange(1, L + 1)
    got = pb.format_rows(pb.KIND_OLS, rec, pos, chr_names=["chr1"], chr_index=np.zeros(L, np.uint32), n_threads=3,
                         exact_p_pools=n).decode()
    spoiled = pb.format_rows(pb.KIND_OLS, rec, pos, chr_names=["chr1"], chr_index=np.zeros(L, np.uint32), n_threads=3).decode()
    expect = "".join(pgo.format_ols_lines("chr1", l + 1, pgo.ols_iterate(counts[l].T.astype(np.uint64), codes, phen, fs))
                     for l in range(L))
    assert got == expect and got.count("\n") > L
    assert spoiled != expect
