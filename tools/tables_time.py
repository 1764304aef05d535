#!/usr/bin/env python
"""Times the count tests on the synthetic C5 shape: usage tables_time.py n_pools n_alleles loci [iters]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb
n, A, L = (int(x) for x in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
ctx = pb.Context(0)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
for kind, name in ((pb.KIND_CHISQ, "chisq_test"), (pb.KIND_FISHER, "fisher_exact_test")):
    scan = pb.Scan(ctx, kind, fs, n, np.arange(A, dtype=np.uint8))
    b = scan.batch(L)
    b.synth(0x5EED0005, 0, L)
    b.time_runs(2)
    ms, _ = b.time_runs(iters)
    per = ms / iters
    print(f"{name} n={n} A={A} L={L}: {per:.3f} ms  {L / per / 1e6:.2f} Gloci/s  {(4 * n * 6 + 16) * L / per / 1e6:.0f} GB/s alg")
    b.close(); scan.close()
ctx.close()
