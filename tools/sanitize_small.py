#!/usr/bin/env python
"""Small invocations of every kernel family in one process (for a debugger or a sanitizer where one is available;
compute-sanitizer is closed on this pool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb

ctx = pb.Context(0)
rng = np.random.default_rng(1)
for n, A, k, L in ((7, 4, 1, 300), (100, 4, 2, 400), (130, 5, 3, 200), (300, 6, 1, 150)):
    counts = pb.synth_counts_host(11, 0, L, n, min(A, 4))
    full = np.zeros((L, A, n), dtype=np.uint32)
    full[:, :min(A, 4)] = counts
    if A > 4:
        full[:, 4:] = rng.random((L, A - 4, n)) < 0.03
    codes = np.arange(A, dtype=np.uint8) if A < 6 else np.arange(6, dtype=np.uint8)
    if A == 5:
        codes = np.array([0, 1, 2, 3, 5], dtype=np.uint8)
    phen = pb.synth_phen_host(11, n, k)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), min_allele_frequency=0.01)
    for kind in (pb.KIND_OLS, pb.KIND_CORR):
        scan = pb.Scan(ctx, kind, fs, n, codes, phen)
        r = scan.run_counts(full)
        r8 = scan.run_counts(full.astype(np.uint8))
        assert np.array_equal(r.stats, r8.stats, equal_nan=True)
        scan.close()
    scan = pb.Scan(ctx, pb.KIND_CHISQ, fs, n, codes)
    scan.run_counts(full)
    scan.close()
    if n <= 16:
        scan = pb.Scan(ctx, pb.KIND_FISHER, fs, n, codes)
        scan.run_counts(full)
        scan.run_counts(full.astype(np.uint16))
        scan.close()
# sync text: synchronous, deferred stream, comment-heavy chunk, kinship loader from text
n, L = 12, 700
buf = np.empty(L * (16 + n * 24), dtype=np.uint8)
nb = pb.synth_sync_text_host(5, 0, L, n, 4, buf)
text = buf[:nb].tobytes()
phen = pb.synth_phen_host(5, n, 2)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, np.arange(6, dtype=np.uint8), phen)
b = scan.batch(L)
nl, off, pos = b.upload_sync_text(text)
b.run()
rec = b.fetch()
rows = pb.format_rows(pb.KIND_OLS, rec, pos, text=text, line_offsets=off)
b.close()
scan.stream_begin(256)
lines = text.split(b"\n")
pend = []
for i in range(0, L, 200):
    chunk = b"\n".join(lines[i:i + 200]) + b"\n"
    if i == 200:
        chunk = b"# c\n" * 6000 + chunk
    pend.append(pb.capi.submit_sync_text(scan, chunk, deferred=True)[0])
    if len(pend) == 3:
        scan.collect(pend.pop(0))
while pend:
    scan.collect(pend.pop(0))
scan.close()
kin = pb.Kinship(ctx, n, 5 * L)
kin.append_sync_text(text, fs, L, True)
kin.gram()
kin.eig_select(kin.columns, 0.75)
kin.covar_scan(phen)
kin.close()
# round 2: Fisher beyond 16 pools, fewer pools than coefficients, the Nelder-Mead analyses, the DMMA covariate scan,
# the kinship n < p branch, the library's communicator with one rank
n, A, L = 20, 4, 200
full = (rng.poisson(0.6, size=(L, A, n)) + (rng.random((L, A, n)) < 0.02) * 40).astype(np.uint32)
full[:, 0] += 1
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), min_allele_frequency=0.0)
scan = pb.Scan(ctx, pb.KIND_FISHER, fs, n, np.arange(A, dtype=np.uint8))
scan.run_counts(full)
scan.close()
n, A, L = 3, 6, 300
full = rng.integers(1, 40, size=(L, A, n)).astype(np.uint32)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n), remove_ns=False, min_allele_frequency=0.0)
pb.ols_iterate(ctx, full, rng.standard_normal((n, 2)), fs, np.arange(A, dtype=np.uint8))
n, A, L = 9, 4, 200
full = pb.synth_counts_host(3, 0, L, n, A)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
pb.mle_iterate(ctx, full, pb.synth_phen_host(3, n, 5), fs, np.arange(A, dtype=np.uint8))
fmt = np.full((n, 3), -np.inf)
fmt[:, 0] = 1.0 / n
fmt[:, 1] = np.linspace(0.0, 0.9, n)
fmt[:3, 2] = (0.1, 0.0, 1.0)
for m in ("LS", "ML"):
    pb.gwalpha(ctx, full[:40], fmt, fs, m, np.arange(A, dtype=np.uint8))
n, P = 37, 203
G = np.clip(0.5 + 0.2 * rng.standard_normal((P, n)), 0, 1)
kin = pb.Kinship(ctx, n, P)
kin.append_columns(G)
kin.set_covariates(rng.standard_normal((n, 6)))
kin.covar_scan(rng.standard_normal((n, 3)))
kin.mle_scan(rng.standard_normal((n, 2)))
kin.gram()
from poolgen_b200 import shard
comm = shard.make_comm(ctx)
comm.kin_allreduce([kin])
kin.eig_select(0, 1.5)      # never reached: n_eigenvecs = n, the n < p branch
kin.covar_scan(rng.standard_normal((n, 2)))
kin.close()
comm.close()
ctx.close()
print("sanitize_small: done", len(rows))
