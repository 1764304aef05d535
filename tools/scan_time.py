#!/usr/bin/env python
"""Times the resident scan kernel for one shape (tuning helper; PG_NBUF / PG_WARPS are read by the library).
usage: scan_time.py n_pools n_alleles k loci [kind] [iters]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb

n, A, k, L = (int(x) for x in sys.argv[1:5])
kind = int(sys.argv[5]) if len(sys.argv) > 5 else pb.KIND_OLS
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 10
ctx = pb.Context(0)
phen = pb.synth_phen_host(0x5EED0003, n, k)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
scan = pb.Scan(ctx, kind, fs, n, np.arange(A, dtype=np.uint8), phen)
b = scan.batch(L)
b.synth(0x5EED0003, 0, L)
b.time_runs(3)
ms, nl = b.time_runs(iters)
per = ms / iters  # one run = the streaming kernel + the fix-up kernel
alg = 8 * n * A + 32 * (A - 1) * k
print(f"n={n} A={A} k={k} L={L} kind={kind} NBUF={os.environ.get('PG_NBUF','-')} WARPS={os.environ.get('PG_WARPS','-')}: "
      f"{per:.3f} ms  {L / per / 1e3:.1f} Mloci/s  {alg * L / per / 1e6:.0f} GB/s alg  frac {alg * L / per / 1e6 / 6551.4:.3f}")
b.close(); scan.close(); ctx.close()
