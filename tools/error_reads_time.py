#!/usr/bin/env python
"""Scan throughput on a real-data-like batch: biallelic loci with stray reads on a third / fourth allele and on D in a
few pools, so that (nearly) every locus needs the renormalised frequencies.  usage: error_reads_time.py n_pools loci"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb
n, L = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(3)
depth = rng.integers(30, 90, (L, n)).astype(np.uint32)
pa = rng.uniform(0.1, 0.9, (L, 1))
a = rng.binomial(depth, np.clip(pa + rng.normal(0, 0.08, (L, n)), 0.02, 0.98)).astype(np.uint32)
full = np.zeros((L, 5, n), dtype=np.uint32)   # A T C G D
full[:, 0] = a
full[:, 1] = depth - a
for col, pr in ((2, 0.02), (3, 0.02), (4, 0.005)):
    full[:, col] = rng.random((L, n)) < pr
ctx = pb.Context(0)
phen = pb.synth_phen_host(3, n, 3)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, np.array([0, 1, 2, 3, 5], dtype=np.uint8), phen)
b = scan.batch(L)
b.upload_counts(full)
b.time_runs(3)
ms, _ = b.time_runs(10)
per = ms / 10
b.download(); b.sync()
rv = b.results_view()
meta = np.ctypeslib.as_array(rv.meta, shape=(L,))
ok = ((meta & 0xFF) == pb.LOCUS_OK).mean()
nout = ((meta >> 8) & 0xFF)[(meta & 0xFF) == pb.LOCUS_OK]
alg = 8 * n * 5 + 32 * 4 * 3
print(f"n={n} L={L} A=5: {per:.3f} ms  {L / per / 1e3:.1f} Mloci/s  {alg * L / per / 1e6:.0f} GB/s alg  frac {alg * L / per / 1e6 / 6551.4:.3f}"
      f"  ok {ok:.3f}  mean rows per locus {nout.mean():.2f}  hints: {os.environ.get('PG_NOHINT', 'on')}")
b.close(); scan.close(); ctx.close()
