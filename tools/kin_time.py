#!/usr/bin/env python
"""Times the kinship path for one shape: usage kin_time.py n_pools n_loci [k] [iters]  (columns = 2 * n_loci)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb

n, L = int(sys.argv[1]), int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = pb.Context(0)
kin = pb.Kinship(ctx, n, 2 * L)
kin.synth(0x5EED0004, 0, L)
P = kin.columns
kin.gram_time(1)
ms = kin.gram_time(iters) / iters
flops = 2.0 * n * n * P
print(f"gram n={n} P={P}: {ms:.2f} ms  algorithmic {flops / ms / 1e9:.2f} TFLOP/s (2 n^2 P), hardware (upper triangle of 128-tiles) "
      f"{flops / ms / 1e9 * (((n + 127) // 128) * ((n + 127) // 128 + 1) / 2 * 128 * 128) / (n * n):.2f} TFLOP/s; frac of 37.1 = {flops / ms / 1e9 / 37.1:.3f}")
t0 = time.perf_counter()
m = kin.eig_select(P, 0.75)
t1 = time.perf_counter()
print(f"eig_select: m={m}, {1e3 * (t1 - t0):.0f} ms, top eigenvalue shares {kin.eigvals(3) / kin.eigvals(n).sum()}")
phen = pb.synth_phen_host(0x5EED0004, n, k)
kin.covar_scan(phen, 1)
_, _, _, ms = kin.covar_scan(phen, iters)
ms /= iters
print(f"covar scan n={n} P={P} k={k} m={m}: {ms:.3f} ms  {P / ms / 1e3:.1f} Mcolumns/s  {(8.0 * n + 24 * k) * P / ms / 1e6:.0f} GB/s alg "
      f"frac {(8.0 * n + 24 * k) * P / ms / 1e6 / 6551.4:.3f}")
kin.close(); ctx.close()
