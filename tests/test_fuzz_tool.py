"""The randomised differential tester (tools/fuzz_parity.py) runs on the GPU box; here: its case generator
is deterministic in the seed (a violation's seed must replay the same case) and covers what it says it covers."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cases_replay_from_their_seed():
    fz = _load()
    kinds, widths, pools, styles = set(), set(), set(), set()
    for seed in range(1000, 1060):
        a = fz.build_case(seed)
        b = fz.build_case(seed)
        assert a[-1] == b[-1]                                   # the label names every drawn parameter
        assert np.array_equal(a[4], b[4]) and np.array_equal(a[7], b[7], equal_nan=True)
        kind, n, codes, k, counts, width, fs, phen, label = a
        assert counts.shape == (counts.shape[0], len(codes), n) and phen.shape == (n, k)
        assert list(codes) == sorted(set(codes)) and 2 <= len(codes) <= 6 and 1 <= k <= 6
        assert width in (8, 16, 32) and counts.max() < 2 ** width
        assert abs(fs.pool_sizes.sum() - 1.0) < 1e-12
        kinds.add(kind)
        widths.add(width)
        pools.add(n)
        styles.add(label.split(" u")[0].split()[-1])
    assert len(kinds) == 4 and len(widths) >= 2 and len(pools) >= 15 and len(styles) == 4
