import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import poolgen_b200 as pb
from tests import helpers as H
from oracle import pgo
import fuzz_parity as fz
from fractions import Fraction as F
import mpmath as mp
mp.mp.dps = 40
ctx = pb.Context(0)
for seed in [int(a) for a in sys.argv[1:]]:
    kind, n, codes, k, counts, width, fs, phen, label = fz.build_case(seed)
    print("==", label)
    ofs = H.oracle_fs(fs)
    scan = pb.Scan(ctx, kind, fs, n, codes, phen)
    dev = scan.run_counts(counts)
    scan.close()
    orc = pgo.scan_batch(pgo.SCAN_OLS if kind == pb.KIND_OLS else pgo.SCAN_CORR, counts, codes, phen, ofs, 8)
    shown = 0
    for l in range(counts.shape[0]):
        if orc.status[l] != pgo.OK or dev.status[l] != pb.LOCUS_OK:
            continue
        m = int(orc.n_out[l])
        if kind == pb.KIND_CORR:
            op, dp = orc.pval[l][:m], dev.stats[l][:m, :, 3]
            with np.errstate(all="ignore"):
                e = np.abs(dp - op) / (np.abs(op) + 2.3e-10)
            bad = np.argwhere(e > 1e-6)
            if bad.size == 0:
                continue
            X = H._design(counts[l], codes, ofs)   # [1 | sorted freqs minus major]?  use raw frequencies instead
            print("locus", l, "counts", counts[l].tolist())
            for s_, j in bad[:2]:
                print("  slot", s_, "phen", j, "r oracle", repr(orc.stat[l][s_, j]), "device raw", repr(dev.stats[l][s_, j, 1]), "p oracle", repr(op[s_, j]), "device", repr(dp[s_, j]))
            # exact r for every kept column against every phenotype
            kept = [a for a in range(len(codes))]
            shown += 1
        else:
            ob, db = orc.stat[l][:m], dev.stats[l][:m, :, 0]
            with np.errstate(all="ignore"):
                e = np.abs(db - ob) / np.maximum(np.abs(ob), 1e-300)
            if not (e > 1e-9).any():
                continue
            X = H._design(counts[l], codes, ofs)
            cond = np.linalg.cond(X.T @ X) if X.shape[0] >= X.shape[1] else np.linalg.cond(X @ X.T)
            if m + 1 < n and shown > 2:
                continue
            print("locus", l, "counts", counts[l].tolist(), "n_out", m, "cond", cond)
            print("  oracle beta", ob.T.tolist(), "\n  device beta", db.T.tolist())
            # exact
            Xf = [[F(float(v)) for v in row] for row in X]
            import sympy
            M = sympy.Matrix(len(Xf), len(Xf[0]), lambda i, j: sympy.Rational(Xf[i][j].numerator, Xf[i][j].denominator))
            for j in range(k):
                yy = sympy.Matrix([sympy.Rational(F(float(v)).numerator, F(float(v)).denominator) for v in phen[:, j]])
                try:
                    bb = (M.T * M).inv() * M.T * yy if M.rows >= M.cols else M.T * (M * M.T).inv() * yy
                    print("  exact phen", j, [float(v) for v in bb][1:])
                except Exception as ex:
                    print("  exact: singular", type(ex).__name__)
            shown += 1
        if shown >= 4:
            break
