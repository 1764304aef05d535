#!/usr/bin/env python
"""Times the Nelder-Mead analyses: usage nm_time.py mle n_pools k loci | nm_time.py gwalpha LS|ML n_pools loci"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb

ctx = pb.Context(0)
A = 4
if sys.argv[1] == "mle":
    n, k, L = (int(v) for v in sys.argv[2:5])
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_MLE, fs, n, np.arange(A, dtype=np.uint8), pb.synth_phen_host(2, n, k))
else:
    n, L = int(sys.argv[3]), int(sys.argv[4])
    rng = np.random.default_rng(1)
    fmt = np.full((max(n, 3), 3), -np.inf)
    fmt[:n, 0] = 1.0 / n
    fmt[:n, 1] = np.concatenate([[0.0], np.sort(rng.uniform(0.05, 0.95, n - 1))])
    fmt[:3, 2] = (0.1, 0.0, 1.0)
    fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
    scan = pb.Scan(ctx, pb.KIND_GWALPHA_LS if sys.argv[2] == "LS" else pb.KIND_GWALPHA_ML, fs, n, np.arange(A, dtype=np.uint8), fmt)
b = scan.batch(L)
b.synth(0x5EED0002, 0, L)
b.time_runs(1)
ms, nl = b.time_runs(1)
print(" ".join(sys.argv[1:]), f": {ms:.1f} ms, {L / ms * 1e3:.0f} loci/s ({nl} launches)")
b.close()
scan.close()
ctx.close()
