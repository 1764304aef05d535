#!/usr/bin/env python
"""Randomised differential test of the per-locus scans against the CPU checker (tests/helpers.py does the comparing):
random pool counts, allele columns, phenotypes, count distributions (sparse, deep, zero-depth pools, monomorphic loci,
near-threshold minor alleles), storage widths and filter settings.  Every case prints one line; a violation prints the
case's seed so that `--only SEED` replays it.

    python tools/fuzz_parity.py --seconds 240 [--seed 1] [--only CASE_SEED] [--mode scan|kin|text]

--mode kin: the column loader (bit-exact columns and labels) and the covariate scan; --mode text: sync text through the
device parser in slabs against the same loci as counts (identical records).
"""
import argparse
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb  # noqa: E402
from tests import helpers as H  # noqa: E402

POOLS = [2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 17, 20, 31, 32, 33, 40, 63, 64, 65, 100, 127, 128, 129, 130, 200, 255, 256, 257,
         300, 511, 512, 513, 1000, 1023, 1025]


def make_counts(rng, L, A, n, style):
    if style == "poisson":
        lam = rng.uniform(0.3, 30.0)
        w = rng.dirichlet(np.full(A, rng.uniform(0.2, 3.0)), size=L)            # allele weights per locus
        c = rng.poisson(lam * A * w[:, :, None] * np.ones((1, 1, n)))
    elif style == "sparse":
        c = (rng.random((L, A, n)) < rng.uniform(0.05, 0.5)) * rng.integers(1, 6, (L, A, n))
        c[:, 0] += rng.integers(0, 3, (L, n))
    elif style == "deep":
        depth = rng.integers(50, 60000, (L, 1, n))
        w = rng.dirichlet(np.full(A, 0.7), size=L)
        c = np.floor(depth * w[:, :, None] * rng.uniform(0.7, 1.3, (L, A, n)))
    else:  # "biallelic": two alleles carry the reads, the minor one near the MAF thresholds
        c = np.zeros((L, A, n))
        a0 = rng.integers(0, A, L)
        a1 = (a0 + rng.integers(1, A, L)) % A
        depth = rng.integers(10, 120, (L, n))
        f = rng.choice([0.0005, 0.001, 0.005, 0.01, 0.05, 0.2, 0.5], size=L)[:, None] * rng.uniform(0.5, 1.5, (L, n))
        minor = rng.binomial(depth, np.clip(f, 0, 1))
        c[np.arange(L), a0] = depth - minor
        c[np.arange(L), a1] = minor
    c = np.asarray(c, dtype=np.int64)
    # zero-depth pools, monomorphic loci
    z = rng.random((L, n)) < rng.choice([0.0, 0.0, 0.002, 0.05])
    c[np.broadcast_to(z[:, None, :], c.shape)] = 0
    mono = rng.random(L) < 0.03
    c[mono, 1:] = 0
    return np.clip(c, 0, 2**31 - 1).astype(np.uint32)


def build_case(seed):
    rng = np.random.default_rng(seed)
    kind = [pb.KIND_OLS, pb.KIND_CORR, pb.KIND_CHISQ, pb.KIND_FISHER][rng.integers(0, 4)]
    n = int(rng.choice(POOLS))
    if kind == pb.KIND_FISHER and n > 40:
        n = int(rng.choice([2, 3, 4, 5, 8, 16, 17, 24, 40]))
    A = int(rng.integers(2, 7))
    codes = np.sort(rng.choice(6, size=A, replace=False)).astype(np.uint8)
    k = int(rng.integers(1, 7))   # beyond the kernels' per-pass limit the scan takes the phenotypes in passes
    L = int(np.clip(rng.integers(100, 3000) * 60 // (n + 20), 40, 3000))
    style = str(rng.choice(["poisson", "sparse", "deep", "biallelic"]))
    if kind == pb.KIND_FISHER and style == "deep":
        style = "poisson"
    counts = make_counts(rng, L, A, n, style)
    width = rng.choice([8, 16, 32])
    if width == 8 and counts.max() > 255 or width == 16 and counts.max() > 65535:
        width = 32
    ps = rng.uniform(1.0, 50.0, n) if rng.random() < 0.5 else np.ones(n)
    tot = 0.0
    for v in ps:
        tot = tot + v
    ps = np.array([v / tot for v in ps])
    fs = pb.FilterStats(pool_sizes=ps, remove_ns=bool(rng.random() < 0.7),
                        min_coverage_depth=int(rng.choice([1, 1, 2, 5, 10, 20])),
                        min_allele_frequency=float(rng.choice([0.0, 0.0005, 0.001, 0.01, 0.05, 0.2])),
                        max_missingness_rate=float(rng.choice([0.0, 0.1, 0.5, 1.0])))
    phen = rng.standard_normal((n, k)) * rng.uniform(0.1, 100.0) + rng.uniform(-50, 50)
    if rng.random() < 0.3:
        phen[:, 0] += 5.0 * counts[min(7, L - 1), 0] / np.maximum(counts[min(7, L - 1)].sum(axis=0), 1)
    if kind == pb.KIND_CORR and n > 6 and rng.random() < 0.3:
        phen[rng.integers(0, n, 2), rng.integers(0, k)] = np.nan
    label = (f"seed={seed} kind={kind} n={n} codes={codes.tolist()} k={k} L={L} {style} u{width} ns={fs.remove_ns} "
             f"depth={fs.min_coverage_depth} maf={fs.min_allele_frequency} miss={fs.max_missingness_rate} "
             f"weighted={bool(ps.max() > ps.min())}")
    return kind, n, codes, k, counts, int(width), fs, phen, label


def one_case(ctx, seed, verbose=True):
    kind, n, codes, k, counts, width, fs, phen, label = build_case(seed)
    up = counts.astype({8: np.uint8, 16: np.uint16, 32: np.uint32}[int(width)])
    print(f"run {label}", flush=True)   # a kernel that never returns leaves this as the last line
    try:
        if kind in (pb.KIND_OLS, pb.KIND_CORR):
            scan = pb.Scan(ctx, kind, fs, n, codes, phen)
            dev = scan.run_counts(up)
            scan.close()
            st = H.compare_regression(kind, counts, codes, phen, fs, dev, label=label)
        else:
            scan = pb.Scan(ctx, kind, fs, n, codes)
            dev = scan.run_counts(up)
            scan.close()
            st = H.compare_tables(kind, counts, codes, fs, dev, label=label)
    except pb.PgError as e:
        if verbose:
            print(f"REFUSED {label}: {e}", flush=True)
        return "refused"
    except AssertionError as e:
        print(f"VIOLATION {label}\n    {e}", flush=True)
        return "violation"
    except Exception:
        print(f"ERROR {label}", flush=True)
        traceback.print_exc()
        return "error"
    if verbose:
        print(f"ok {label} -> {st}", flush=True)
    return "ok"


def kin_case(ctx, seed, verbose=True):
    """LoadAll on the device (columns and labels bit-exact) and the covariate scan with 0..4 explicit covariates"""
    from oracle import pgo
    from tests.test_kinship_gpu import _cmp_records
    rng = np.random.default_rng(seed)
    n = int(rng.choice([3, 4, 5, 6, 7, 9, 15, 16, 17, 24, 31, 33, 48, 64, 65, 100, 130, 300]))
    A = int(rng.integers(2, 7))
    codes = np.sort(rng.choice(6, size=A, replace=False)).astype(np.uint8)
    L = int(np.clip(rng.integers(30, 400) * 40 // (n + 20), 10, 300))
    style = str(rng.choice(["poisson", "sparse", "deep", "biallelic"]))
    counts = make_counts(rng, L, A, n, style)
    ps = rng.uniform(1.0, 50.0, n) if rng.random() < 0.5 else np.ones(n)
    tot = 0.0
    for v in ps:
        tot = tot + v
    ps = np.array([v / tot for v in ps])
    fs = pb.FilterStats(pool_sizes=ps, remove_ns=bool(rng.random() < 0.7),
                        min_coverage_depth=int(rng.choice([1, 1, 2, 5, 10])),
                        min_allele_frequency=float(rng.choice([0.0, 0.0005, 0.001, 0.01, 0.05])),
                        max_missingness_rate=float(rng.choice([0.0, 0.1, 1.0])))
    keep = bool(rng.random() < 0.5)
    m = int(rng.integers(0, min(5, n - 2)))
    k = int(rng.integers(1, 4))
    cov = rng.standard_normal((n, m))
    phen = rng.standard_normal((n, k)) * rng.uniform(0.1, 30.0) + rng.uniform(-20, 20)
    label = (f"kin seed={seed} n={n} codes={codes.tolist()} L={L} {style} ns={fs.remove_ns} depth={fs.min_coverage_depth} "
             f"maf={fs.min_allele_frequency} miss={fs.max_missingness_rate} keep_p_minus_1={keep} m={m} k={k}")
    print(f"run {label}", flush=True)
    try:
        kin = pb.Kinship(ctx, n, 6 * L)
        loc, alle = kin.append_counts(counts, codes, fs, keep)
        P = kin.columns
        G = kin.get_columns(0, P) if P else np.zeros((0, n))
        ocols, olabels = pgo.load_columns(counts.transpose(0, 2, 1).astype(np.uint64), codes, H.oracle_fs(fs), keep)
        assert len(olabels) == P, f"{label}: {P} columns, oracle {len(olabels)}"
        assert [l for l, _ in olabels] == list(loc) and [a for _, a in olabels] == list(alle), f"{label}: labels differ"
        assert P == 0 or np.array_equal(G, ocols, equal_nan=True), f"{label}: columns differ"
        narb = 0
        if P and not np.isnan(G).any():
            kin.set_covariates(cov)
            beta, var, pval = kin.covar_scan(phen)
            ob, ov, op = (np.full((P, k), np.nan) for _ in range(3))
            for c in range(P):
                x = np.ones((n, 2 + m))
                x[:, 1:1 + m] = cov
                x[:, 1 + m] = G[c]
                rc, b, v, p_, _ = pgo.ols(x, phen)
                if rc == 0:
                    ob[c], ov[c], op[c] = b[1 + m], v[1 + m], p_[1 + m]
            # columns inside span[1 | covariates] (constant over the pools): NaN on the device, rounding noise in the oracle
            flat = np.ptp(G, axis=1) == 0
            ob[flat], ov[flat], op[flat] = np.nan, np.nan, np.nan
            beta[:, flat], var[:, flat], pval[:, flat] = np.nan, np.nan, np.nan
            narb = _cmp_records((beta, var, pval), (ob.T, ov.T, op.T), label, arb=(G, cov, phen))
        kin.close()
    except pb.PgError as e:
        if verbose:
            print(f"REFUSED {label}: {e}", flush=True)
        return "refused"
    except AssertionError as e:
        print(f"VIOLATION {label}\n    {str(e)[:600]}", flush=True)
        return "violation"
    except Exception:
        print(f"ERROR {label}", flush=True)
        traceback.print_exc()
        return "error"
    if verbose:
        print(f"ok {label} -> {P} columns, {narb} arbitrated", flush=True)
    return "ok"


def text_case(ctx, seed, verbose=True):
    """the same loci as sync text through the device parser (one submit per slab, random slab sizes) and as counts:
    identical records"""
    from tests.test_text_gpu import _sync_text
    rng = np.random.default_rng(seed)
    n = int(rng.choice([2, 3, 5, 8, 16, 17, 40, 100, 130, 300]))
    L = int(np.clip(rng.integers(50, 1500) * 40 // (n + 20), 20, 1500))
    k = int(rng.integers(1, 4))
    kind = [pb.KIND_OLS, pb.KIND_CORR, pb.KIND_CHISQ, pb.KIND_FISHER][rng.integers(0, 4)]
    if kind == pb.KIND_FISHER and n > 40:
        n = 17
    style = str(rng.choice(["poisson", "sparse", "deep", "biallelic"]))
    if kind == pb.KIND_FISHER and style == "deep":
        style = "poisson"
    counts = make_counts(rng, L, 6, n, style)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), remove_ns=bool(rng.random() < 0.7),
                        min_coverage_depth=int(rng.choice([1, 2, 10])),
                        min_allele_frequency=float(rng.choice([0.0, 0.001, 0.01, 0.05])),
                        max_missingness_rate=float(rng.choice([0.0, 0.1, 1.0])))
    phen = rng.standard_normal((n, k)) if kind in (pb.KIND_OLS, pb.KIND_CORR) else None
    codes = np.arange(6, dtype=np.uint8)
    slab = int(rng.choice([7, 64, 333, 1000, 4096]))
    crlf = bool(rng.random() < 0.3)
    label = f"text seed={seed} kind={kind} n={n} L={L} k={k} {style} slab={slab} crlf={crlf} ns={fs.remove_ns}"
    print(f"run {label}", flush=True)
    try:
        scan = pb.Scan(ctx, kind, fs, n, codes, phen)
        whole = scan.run_counts(counts)
        scan.stream_begin(slab)
        chroms = ["chr%d" % (1 + l // 400) for l in range(L)]
        pos = [10 + 3 * l for l in range(L)]
        pending, parts = [], []
        for l0 in range(0, L, slab):
            text = _sync_text(counts[l0:l0 + slab], chroms[l0:l0 + slab], pos[l0:l0 + slab], crlf)
            t, nl = pb.capi.submit_sync_text(scan, text)
            assert nl == min(slab, L - l0), f"{label}: parsed {nl} loci"
            pending.append(t)
            if len(pending) == 2:
                parts.append(scan.collect(pending.pop(0)))
        while pending:
            parts.append(scan.collect(pending.pop(0)))
        scan.close()
        status = np.concatenate([p.status for p in parts])
        stats = np.concatenate([p.stats for p in parts])
        assert (status == whole.status).all(), f"{label}: status differs"
        assert np.array_equal(stats, whole.stats, equal_nan=True), f"{label}: records differ"
    except pb.PgError as e:
        if verbose:
            print(f"REFUSED {label}: {e}", flush=True)
        return "refused"
    except AssertionError as e:
        print(f"VIOLATION {label}\n    {str(e)[:600]}", flush=True)
        return "violation"
    except Exception:
        print(f"ERROR {label}", flush=True)
        traceback.print_exc()
        return "error"
    if verbose:
        print(f"ok {label}", flush=True)
    return "ok"


def nm_case(ctx, seed, verbose=True):
    """mle_iter and gwalpha: keep-mask, status, allele order exact; values to the solver's convergence (loose here: the
    point is that every shape runs to the end with the statuses of the CPU checker)"""
    from oracle import pgo
    rng = np.random.default_rng(seed)
    which = str(rng.choice(["mle", "gw_ls", "gw_ml"]))
    n = int(rng.choice([2, 3, 4, 5, 6, 8, 12, 16, 31, 33, 64] if which != "mle" else [2, 3, 4, 5, 6, 9, 17, 33, 100, 130, 300]))
    A = int(rng.integers(2, 7))
    codes = np.sort(rng.choice(6, size=A, replace=False)).astype(np.uint8)
    k = int(rng.integers(1, 5)) if which == "mle" else 1
    L = int(rng.integers(20, 160)) if which != "mle" else int(rng.integers(50, 500))
    style = str(rng.choice(["poisson", "sparse", "deep", "biallelic"]))
    counts = make_counts(rng, L, A, n, style)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), remove_ns=bool(rng.random() < 0.7),
                        min_coverage_depth=int(rng.choice([1, 2, 10])),
                        min_allele_frequency=float(rng.choice([0.0, 0.001, 0.01, 0.05])),
                        max_missingness_rate=float(rng.choice([0.0, 0.1, 1.0])))
    label = f"nm seed={seed} {which} n={n} codes={codes.tolist()} k={k} L={L} {style} ns={fs.remove_ns} depth={fs.min_coverage_depth} maf={fs.min_allele_frequency}"
    print(f"run {label}", flush=True)
    try:
        if which == "mle":
            phen = rng.standard_normal((n, k)) * rng.uniform(0.5, 10.0) + rng.uniform(-5, 5)
            dev = pb.mle_iterate(ctx, counts, phen, fs, codes)
            orc = pgo.scan_batch(pgo.SCAN_MLE, counts, codes, phen, H.oracle_fs(fs), 8)
        else:
            bins = rng.dirichlet(np.full(n, 8.0))
            q = np.concatenate([[0.0], np.sort(rng.uniform(0.05, 0.95, size=n - 1))])
            rows = max(n, 3)
            phen = np.full((rows, 3), -np.inf)
            phen[:n, 0], phen[:n, 1] = bins, q
            phen[:3, 2] = (0.15, 0.0, 1.0)
            dev = pb.gwalpha(ctx, counts, phen, fs, "LS" if which == "gw_ls" else "ML", codes)
            orc = pgo.scan_batch(pgo.SCAN_GWALPHA_LS if which == "gw_ls" else pgo.SCAN_GWALPHA_ML, counts, codes, phen,
                                 H.oracle_fs(fs), 8)
        assert ((orc.status == pgo.FILTERED) == (dev.status == pb.LOCUS_FILTERED)).all(), f"{label}: keep-mask differs"
        both = (orc.status == pgo.OK) & (dev.status == pb.LOCUS_OK)
        agree = ((orc.status == pgo.OK) == (dev.status == pb.LOCUS_OK)) | (dev.status == pb.LOCUS_UNSUPPORTED)
        assert agree.mean() > 0.97, f"{label}: status differs at {np.nonzero(~agree)[0][:8]} oracle {orc.status[~agree][:8]} device {dev.status[~agree][:8]}"
        assert (orc.n_out[both] == dev.n_out[both]).all() and (orc.allele[both] == dev.alleles[both]).all(), f"{label}: order differs"
        med = 0.0
        if both.any():
            S = dev.stats.shape[1]
            kk = dev.stats.shape[2]
            slot = np.broadcast_to((np.arange(S)[None, :] < orc.n_out[both][:, None])[:, :, None], (both.sum(), S, kk))
            o = orc.stat[both][:, :S].reshape(both.sum(), S, -1)[:, :, :kk][slot]
            d = dev.stats[both][..., 0][slot]
            fin = np.isfinite(o) & np.isfinite(d)
            assert (np.isfinite(o) == np.isfinite(d)).mean() > 0.97, f"{label}: NaN pattern differs"
            if fin.any():
                err = np.abs(d[fin] - o[fin]) / np.maximum(np.abs(o[fin]), 1.0)
                med = float(np.median(err))
                assert med < 1e-3, f"{label}: median error {med}"
    except pb.PgError as e:
        if verbose:
            print(f"REFUSED {label}: {e}", flush=True)
        return "refused"
    except AssertionError as e:
        print(f"VIOLATION {label}\n    {str(e)[:600]}", flush=True)
        return "violation"
    except Exception:
        print(f"ERROR {label}", flush=True)
        traceback.print_exc()
        return "error"
    if verbose:
        print(f"ok {label} median err {med:.2e}", flush=True)
    return "ok"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--only", type=int, default=None)
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--mode", choices=["scan", "kin", "text", "nm"], default="scan")
    a = ap.parse_args()
    ctx = pb.Context(0)
    case = {"scan": one_case, "kin": kin_case, "text": text_case, "nm": nm_case}[a.mode]
    if a.only is not None:
        print(case(ctx, a.only))
        return 0
    t0 = time.time()
    tally = {}
    i = 0
    while time.time() - t0 < a.seconds:
        r = case(ctx, a.seed * 1_000_003 + i, verbose=not a.quiet)
        tally[r] = tally.get(r, 0) + 1
        i += 1
    print("fuzz_parity:", tally, flush=True)
    ctx.close()
    return 1 if tally.get("violation", 0) or tally.get("error", 0) else 0


if __name__ == "__main__":
    sys.exit(main())
