#!/usr/bin/env python
"""Times the sync-text path for one shape (tuning helper): pinned text -> H2D -> device parse -> ingest -> scan -> D2H.
usage: text_time.py n_pools slab_loci n_slabs"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb

n, slab, n_slabs = (int(x) for x in sys.argv[1:4])
ctx = pb.Context(0)
lib = pb.capi.lib()
host, hptr = ctx.pinned_empty((2, slab * (16 + n * 24)), np.uint8)
nb = [pb.synth_sync_text_host(0x5EED0003, i * slab, slab, n, 4, host[i]) for i in range(2)]
phen = pb.synth_phen_host(0x5EED0003, n, 3)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
scan = pb.Scan(ctx, pb.KIND_OLS, fs, n, np.arange(6, dtype=np.uint8), phen)
scan.stream_begin(slab)


def run(count):
    pending = []
    for i in range(count):
        t = C.c_int()
        rc = lib.pg_scan_submit_sync_text(scan._h, host[i % 2].ctypes.data, nb[i % 2], C.byref(t), None)
        assert rc == 0, rc
        pending.append(t.value)
        if len(pending) == 3:
            assert scan.collect(pending.pop(0), copy=False).n_loci == slab
    while pending:
        assert scan.collect(pending.pop(0), copy=False).n_loci == slab


run(3)
t0 = time.perf_counter()
run(n_slabs)
dt = time.perf_counter() - t0
print(f"n={n} slab={slab} x {n_slabs}: {dt / n_slabs * 1e3:.3f} ms per slab, {slab * n_slabs / dt / 1e6:.3f} Mloci/s, "
      f"{nb[0] * n_slabs / dt / 1e9:.2f} GB/s of text ({nb[0] / slab:.0f} B per locus)")
scan.close()
ctx.pinned_free(hptr)
ctx.close()

# the box's H2D wire rate for the same bytes (pinned torch tensor), for comparison
import torch
src = torch.empty(int(nb[0]), dtype=torch.uint8).pin_memory()
dst = torch.empty(int(nb[0]), dtype=torch.uint8, device="cuda")
for _ in range(3):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(12):
    dst.copy_(src, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print(f"wire: {e0.elapsed_time(e1) / 12:.3f} ms per {nb[0] / 1e6:.1f} MB = {nb[0] * 12 / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
