"""Sync text parsed on the device (pg_batch_upload_sync_text) against the count path and the oracle's line parser."""
import numpy as np
import pytest

import poolgen_b200 as pb
from oracle import pgo
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _sync_text(counts, chroms, positions, crlf=False, extra=()):
    """counts [L, 6, n] -> sync text (src/base/sync.rs:100-156 format); `extra` = (line index, raw line) insertions"""
    eol = "\r\n" if crlf else "\n"
    lines = ["#chr\tpos\tref\t" + "\t".join(f"pool{i}" for i in range(counts.shape[2]))]
    ins = dict(extra)
    for l in range(counts.shape[0]):
        if l in ins:
            lines.append(ins[l])
        pools = "\t".join(":".join(str(int(v)) for v in counts[l, :, i]) for i in range(counts.shape[2]))
        lines.append(f"{chroms[l]}\t{positions[l]}\tN\t{pools}")
    return (eol.join(lines) + eol).encode()


@pytest.mark.parametrize("crlf", [False, True])
def test_c1_text_matches_counts(ctx, crlf):
    c1 = H.load_c1()
    counts = c1["counts"]
    L = counts.shape[0]
    chroms = [str(c1["chrom_names"][i]) for i in c1["chrom_idx"]]
    pos = [int(p) for p in c1["pos"]]
    text = _sync_text(counts, chroms, pos, crlf, extra=[(5, "Chromosome1\tnot_a_number\tC\t" + "\t".join(["1:2:3:4:5:6"] * 5)),
                                                       (9, "#a comment in the middle")])
    # the oracle's restatement of lparse agrees with the generated text line by line
    body = [ln for ln in text.decode().replace("\r\n", "\n").split("\n") if ln and not ln.startswith("#") and "not_a_number" not in ln]
    for l in (0, 1, 77, L - 1):
        n, ch, p, oc = pgo.parse_sync_line(body[l] + "\n")
        assert n == 5 and ch == chroms[l] and p == pos[l] and (oc.T == counts[l]).all()
    fs = pb.FilterStats(pool_sizes=c1["pool_sizes"])
    for kind, phen in ((pb.KIND_CHISQ, None), (pb.KIND_FISHER, None), (pb.KIND_OLS, c1["phen"]), (pb.KIND_CORR, c1["phen"])):
        scan = pb.Scan(ctx, kind, fs, 5, c1["codes"], phen)
        b = scan.batch(L)
        nl, off, p = b.upload_sync_text(text)
        assert nl == L
        assert list(p) == pos
        for l in (0, 5, 6, 9, 10, L - 1):  # the line offsets point at the chromosome names
            assert text[int(off[l]):].startswith((chroms[l] + "\t" + str(pos[l])).encode())
        b.run()
        from_text = b.fetch()
        b.upload_counts(counts)
        b.run()
        from_counts = b.fetch()
        b.close()
        scan.close()
        assert (from_text.status == from_counts.status).all()
        assert np.array_equal(from_text.stats, from_counts.stats, equal_nan=True)
        assert (from_text.alleles == from_counts.alleles).all()


def test_synthetic_text_and_streaming(ctx):
    n, L, k = 100, 3000, 2
    counts4 = pb.synth_counts_host(0x7E47, 0, L, n, 4)
    counts = np.zeros((L, 6, n), dtype=np.uint32)
    counts[:, :4] = counts4
    counts[:, 5] = (np.arange(L)[:, None] + np.arange(n)[None, :]) % 3 == 0  # a few deletions
    chroms = ["chr%d" % (1 + l // 1000) for l in range(L)]
    pos = [1000 + 7 * l for l in range(L)]
    phen = pb.synth_phen_host(0x7E47, n, k)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    codes = np.arange(6, dtype=np.uint8)
    scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, codes, phen)
    whole = scan.run_counts(counts)
    scan.stream_begin(1024)
    parts = []
    pending = []
    for l0 in range(0, L, 1000):
        text = _sync_text(counts[l0:l0 + 1000], chroms[l0:l0 + 1000], pos[l0:l0 + 1000])
        t, nl = pb.capi.submit_sync_text(scan, text)
        assert nl == min(1000, L - l0)
        pending.append(t)
    for t in pending:
        parts.append(scan.collect(t))
    scan.close()
    status = np.concatenate([p.status for p in parts])
    stats = np.concatenate([p.stats for p in parts])
    assert (status == whole.status).all()
    assert np.array_equal(stats, whole.stats, equal_nan=True)
    H.compare_regression(pb.KIND_OLS, counts, codes, phen, fs, whole, label="synthetic text")


def test_text_errors(ctx):
    n = 3
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_CHISQ, fs, n, np.arange(6, dtype=np.uint8))
    b = scan.batch(16)
    good = b"c\t1\tA\t1:2:3:4:0:0\t1:2:3:4:0:0\t4:3:2:1:0:0\n"
    assert b.upload_sync_text(good)[0] == 1
    assert b.upload_sync_text(good[:-1])[0] == 1                      # last line without its newline
    assert b.upload_sync_text(b"# only a comment\n")[0] == 0
    assert b.upload_sync_text(b"")[0] == 0
    assert b.upload_sync_text(b"c\t+12\tA\t1:2:3:4:0:0:9\t1:2:3:4:0:0\t4:3:2:1:0:0\n")[0] == 1   # extra numbers ignored
    with pytest.raises(pb.PgError):   # two pools where the scan has three
        b.upload_sync_text(b"c\t1\tA\t1:2:3:4:0:0\t1:2:3:4:0:0\n")
    with pytest.raises(pb.PgError):   # five numbers in a pool field
        b.upload_sync_text(b"c\t1\tA\t1:2:3:4:0\t1:2:3:4:0:0\t4:3:2:1:0:0\n")
    with pytest.raises(pb.PgError):   # not an integer
        b.upload_sync_text(b"c\t1\tA\t1:2:x:4:0:0\t1:2:3:4:0:0\t4:3:2:1:0:0\n")
    with pytest.raises(pb.PgError):   # more loci than the batch holds
        b.upload_sync_text(good * 17)
    b.close()
    scan.close()
    scan4 = pb.Scan(ctx, pb.KIND_CHISQ, fs, n, np.arange(4, dtype=np.uint8))
    b4 = scan4.batch(4)
    with pytest.raises(pb.PgError):   # sync text needs the six sync columns
        b4.upload_sync_text(good)
    b4.close()
    scan4.close()


def test_host_text_generator_round_trip(ctx):
    """pg_synth_sync_text_host writes the synthetic counts as sync text: parsing it on the device gives the records of
    the count path"""
    n, A, L = 12, 4, 500
    out = np.empty(L * (16 + n * 24), dtype=np.uint8)
    nb = pb.synth_sync_text_host(0xBEEF, 10, L, n, A, out)
    text = out[:nb].tobytes()
    assert text.count(b"\n") == L and text.startswith(b"chr1\t11\tN\t")
    counts4 = pb.synth_counts_host(0xBEEF, 10, L, n, A)
    counts = np.zeros((L, 6, n), dtype=np.uint32)
    counts[:, :A] = counts4
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_FISHER, fs, n, np.arange(6, dtype=np.uint8))
    b = scan.batch(L)
    nl, off, pos = b.upload_sync_text(text)
    assert nl == L and list(pos) == list(range(11, 11 + L))
    b.run()
    rt = b.fetch()
    b.upload_counts(counts)
    b.run()
    rc = b.fetch()
    b.close()
    scan.close()
    assert (rt.status == rc.status).all() and np.array_equal(rt.stats, rc.stats, equal_nan=True)


@pytest.mark.parametrize("kind", [pb.KIND_OLS, pb.KIND_CORR, pb.KIND_CHISQ, pb.KIND_FISHER])
def test_c1_text_in_csv_out(ctx, kind):
    """the whole replaced stretch of `read_analyse_write` (src/base/sync.rs:788-970): sync text -> device parse -> scan
    -> records -> pg_format_rows, against the lines the oracle's callbacks format for the same file.  Labels, row order
    and row count are exact; the numbers agree within the parity tolerances
    (a printed digit can differ where the device value sits within 1e-9 of a rounding boundary)."""
    c1 = H.load_c1()
    counts = c1["counts"]
    L = counts.shape[0]
    names = [str(s) for s in c1["chrom_names"]]
    chroms = [names[i] for i in c1["chrom_idx"]]
    pos = [int(p) for p in c1["pos"]]
    text = _sync_text(counts, chroms, pos)
    regression = kind in (pb.KIND_OLS, pb.KIND_CORR)
    phen = c1["phen"] if regression else None
    fs = pb.FilterStats(pool_sizes=c1["pool_sizes"])
    scan = pb.Scan(ctx, kind, fs, 5, c1["codes"], phen)
    b = scan.batch(L)
    nl, off, p = b.upload_sync_text(text)
    b.run()
    rec = b.fetch()
    b.close()
    scan.close()
    ofs = H.oracle_fs(fs)
    expect = []
    skipped = 0
    for l in range(L):
        c = counts[l].T.astype(np.uint64)
        if kind == pb.KIND_OLS:
            res = pgo.ols_iterate(c, c1["codes"], phen, ofs)
            line = pgo.format_ols_lines(chroms[l], pos[l], res)
        elif kind == pb.KIND_CORR:
            res = pgo.correlation(c, c1["codes"], phen, ofs)
            line = pgo.format_corr_lines(chroms[l], pos[l], res)
        elif kind == pb.KIND_CHISQ:
            res = pgo.chisq(c, c1["codes"], ofs)
            line = pgo.format_chisq_line(chroms[l], pos[l], res)
        else:
            res = pgo.fisher(c, c1["codes"], ofs)
            line = pgo.format_fisher_line(chroms[l], pos[l], res)
        if (res.status == pgo.OK) != (rec.status[l] == pb.LOCUS_OK):
            # a rank-deficient design (five pools, up to four regressors): one side's solve succeeds, the other's
            # fails -- tests/helpers.py:compare_regression checks these loci are singular; no row comparison
            assert kind == pb.KIND_OLS and res.status != pgo.FILTERED and rec.status[l] != pb.LOCUS_FILTERED
            rec.status[l] = pb.LOCUS_FAILED
            skipped += 1
            continue
        expect.append(line)
    assert skipped < 0.02 * L
    expect = "".join(expect)
    got = pb.format_rows(kind, rec, p, text=text, line_offsets=off, n_threads=4, exact_p_pools=5).decode()
    gl, el = got.strip().split("\n"), expect.strip().split("\n")
    assert len(gl) == len(el) > 6000
    n_text = 3 if not regression else 3
    same = ill = 0
    locus_of = {(chroms[l], str(pos[l])): l for l in range(L)}
    for g, e in zip(gl, el):
        if g == e:
            same += 1
            continue
        gf, ef = g.split(","), e.split(",")
        assert len(gf) == len(ef)
        close = True
        for i, (a, b_) in enumerate(zip(gf, ef)):
            if i < n_text or a.startswith("Pheno_"):
                assert a == b_, (g, e)
            else:
                x, y = float(a), float(b_)
                close &= (x != x and y != y) or abs(x - y) <= 2e-6 * max(abs(x), abs(y)) + 1.1e-6
        if not close:
            # only a (nearly) rank-deficient or saturated design may differ beyond the tolerance (five pools, up to
            # four regressors); tests/helpers.py:compare_regression arbitrates those with a 50-digit solve
            assert kind == pb.KIND_OLS, (g, e)
            X = H._design(counts[locus_of[(gf[0], gf[1])]], c1["codes"], ofs)
            assert X.shape[1] >= X.shape[0] or np.linalg.cond(X.T @ X) > 1e8, (g, e)
            ill += 1
    # ols_iter prints rounded numbers (8 / 6 / 12 digits): nearly every row is the same text; the other analyses print
    # 17 significant digits of the mean frequency and of p, where a last-bit difference shows
    # with PG_FORMAT_EXACT_P the p-value digits are the reference's own for the device's t
    assert ill <= 0.01 * len(el) and (kind != pb.KIND_OLS or same >= 0.95 * len(el)), (same, ill, len(el))
    print(f"kind {kind}: {same} of {len(el)} rows identical text")


def test_deferred_text_stream(ctx):
    """pg_scan_submit_sync_text with n_loci = NULL: the copy and the parse of slab i+1 are enqueued before the host
    waits for slab i; records, labels and row text equal the synchronous path; a chunk with more comment lines than
    the default line bound is re-parsed with the exact bound; errors surface from the call that finishes the slab."""
    n, L, k = 64, 2500, 1
    counts4 = pb.synth_counts_host(0xD3F, 0, L, n, 4)
    counts = np.zeros((L, 6, n), dtype=np.uint32)
    counts[:, :4] = counts4
    chroms = ["scaf%d" % (l // 700) for l in range(L)]
    pos = [5 + 3 * l for l in range(L)]
    phen = pb.synth_phen_host(0xD3F, n, k)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    codes = np.arange(6, dtype=np.uint8)
    scan = pb.Scan(ctx, pb.KIND_CORR, fs, n, codes, phen)
    whole = scan.run_counts(counts)
    scan.stream_begin(600)
    chunks = []
    for l0 in range(0, L, 500):
        extra = [(i, "# filler %d" % i) for i in range(0, 500, 1)] * 12 if l0 == 1000 else ()
        if extra:  # 6,000 comment lines in one chunk: more than capacity + capacity / 4 + 4,096
            body = _sync_text(counts[l0:l0 + 500], chroms[l0:l0 + 500], pos[l0:l0 + 500]).decode().split("\n")
            body = body[:1] + ["# filler"] * 6000 + body[1:]
            chunks.append("\n".join(body).encode())
        else:
            chunks.append(_sync_text(counts[l0:l0 + 500], chroms[l0:l0 + 500], pos[l0:l0 + 500]))
    pending, parts, rows = [], [], []

    def finish(item):
        t, text = item
        rec = scan.collect(t)
        off, p = pb.capi.text_labels(scan, t, rec.status.shape[0])
        parts.append(rec)
        rows.append(pb.format_rows(pb.KIND_CORR, rec, p, text=text, line_offsets=off, n_threads=2))
        for l in (0, rec.status.shape[0] - 1):
            assert text[int(off[l]):].startswith(b"scaf")

    for text in chunks:
        t, nl = pb.capi.submit_sync_text(scan, text, deferred=True)
        assert nl is None
        pending.append((t, text))
        if len(pending) == 3:
            finish(pending.pop(0))
    while pending:
        finish(pending.pop(0))
    status = np.concatenate([p.status for p in parts])
    stats = np.concatenate([p.stats for p in parts])
    assert status.shape[0] == L and (status == whole.status).all()
    assert np.array_equal(stats, whole.stats, equal_nan=True)
    expect = pb.format_rows(pb.KIND_CORR, whole, pos, chr_names=sorted(set(chroms), key=chroms.index),
                            chr_index=[l // 700 for l in range(L)], n_threads=1)
    assert b"".join(rows) == expect and expect.count(b"\n") > 2000
    # a malformed slab in deferred mode: the error comes from the call that finishes it
    bad = b"c\t1\tA\t" + b"\t".join([b"1:2:3:4:0:0"] * (n - 1)) + b"\n"
    t, _ = pb.capi.submit_sync_text(scan, bad, deferred=True)
    with pytest.raises(pb.PgError):
        scan.collect(t)
    scan.close()


def test_text_fuzz_against_the_oracle_parser(ctx):
    """random chunks (comment lines, CRLF, '+' signs, non-integer positions, extra numbers in a pool field, a last line
    without its newline): the loci, positions and line offsets the device parser reports equal what the oracle's
    restatement of `lparse` keeps, and the records equal those of the count path fed with the oracle-parsed counts"""
    rng = np.random.default_rng(20261018)
    n = 7
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), min_allele_frequency=0.01)
    scan = pb.Scan(ctx, pb.KIND_CHISQ, fs, n, np.arange(6, dtype=np.uint8))
    b = scan.batch(256)
    for trial in range(40):
        crlf = bool(rng.integers(0, 2))
        eol = "\r\n" if crlf else "\n"
        lines = []
        for l in range(int(rng.integers(1, 200))):
            r = rng.random()
            if r < 0.08:
                lines.append("#" + "x" * int(rng.integers(0, 30)))
                continue
            pos = str(int(rng.integers(0, 2 ** 40)))
            if r < 0.14:
                pos = rng.choice(["1e5", "12a", "-3", "", "0x10"])
            elif r < 0.2:
                pos = "+" + pos
            fields = []
            for i in range(n):
                c = [str(int(v)) for v in rng.integers(0, 60, 6) * (rng.random(6) < 0.6)]
                if rng.random() < 0.05:
                    c[int(rng.integers(0, 6))] = "+" + c[0]
                if rng.random() < 0.05:
                    c.append("17")
                fields.append(":".join(c))
            lines.append(f"chr{int(rng.integers(1, 4))}\t{pos}\tN\t" + "\t".join(fields))
        text = eol.join(lines) + (eol if rng.random() < 0.7 else "")
        raw = text.encode()
        kept, offs, cnts = [], [], []
        cursor = 0
        for ln in text.split(eol) if text else []:
            start = cursor
            cursor += len(ln.encode()) + len(eol)
            if ln == "" or ln.startswith("#"):
                continue
            nn, ch, p, oc = pgo.parse_sync_line(ln + "\n")
            if nn <= 0:
                continue
            assert nn == n
            kept.append(p)
            offs.append(start)
            cnts.append(oc.T)
        nl, off, p = b.upload_sync_text(raw)
        assert nl == len(kept), (trial, nl, len(kept))
        assert list(p) == kept and list(off) == offs
        if nl == 0:
            continue
        b.run()
        rt = b.fetch()
        b.upload_counts(np.ascontiguousarray(np.array(cnts, dtype=np.uint32)))
        b.run()
        rc = b.fetch()
        assert (rt.status == rc.status).all() and np.array_equal(rt.stats, rc.stats, equal_nan=True)
    b.close()
    scan.close()
