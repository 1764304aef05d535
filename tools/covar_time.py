#!/usr/bin/env python
"""Times the covariate scan with m explicit covariates: usage covar_time.py n_pools n_loci k m [iters]
(columns = 2 * n_loci; PG_COVAR_NO_MMA=1 forces the per-warp dot-product kernels, PG_CM_WARPS=w the warps per CTA of
the DMMA kernel)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb

n, L, k, m = (int(v) for v in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
ctx = pb.Context(0)
kin = pb.Kinship(ctx, n, 2 * L)
kin.synth(0x5EED0004, 0, L)
P = kin.columns
phen = pb.synth_phen_host(0x5EED0004, n, k)
rng = np.random.default_rng(0x5EED0004)
kin.set_covariates(rng.standard_normal((n, m)))
kin.covar_scan(phen, 1)
*_, ms = kin.covar_scan(phen, iters)
ms /= iters
alg = (8.0 * n + 24 * k) * P
fp64_ms = 2.0 * n * (1 + m + k + 1) * P / 37.1e9
hbm_ms = alg / 6551.4e6
print(f"covar scan n={n} P={P} k={k} m={m} ({os.environ.get('PG_COVAR_NO_MMA') and 'dot products' or 'DMMA'}, "
      f"warps {os.environ.get('PG_CM_WARPS', 'auto')}): {ms:.3f} ms  {alg / ms / 1e6:.0f} GB/s alg = {alg / ms / 1e6 / 6551.4:.3f} of HBM; "
      f"floors: HBM {hbm_ms:.2f} ms, FP64 {fp64_ms:.2f} ms -> {max(hbm_ms, fp64_ms) / ms:.3f} of min(HBM, FP64)")
kin.close()
ctx.close()
