#!/usr/bin/env python
"""Sync-text path with T reader threads (one scan handle each), repeated: usage text_threads.py n_pools slab n_slabs T reps"""
import ctypes as C, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poolgen_b200 as pb
n, slab, n_slabs, T, reps = (int(x) for x in sys.argv[1:6])
ctx = pb.Context(0)
lib = pb.capi.lib()
host, hptr = ctx.pinned_empty((2, slab * (16 + n * 24)), np.uint8)
nb = [pb.synth_sync_text_host(0x5EED0003, i * slab, slab, n, 4, host[i]) for i in range(2)]
phen = pb.synth_phen_host(0x5EED0003, n, 3)
fs = pb.FilterStats(pool_sizes=np.full(n, 1.0 / n))
scans = [pb.Scan(ctx, pb.KIND_OLS, fs, n, np.arange(6, dtype=np.uint8), phen) for _ in range(T)]
for s in scans:
    s.stream_begin(slab)
def reader(t, count):
    sc = scans[t]; pending = []
    for i in range(count):
        tk = C.c_int()
        assert lib.pg_scan_submit_sync_text(sc._h, host[i % 2].ctypes.data, nb[i % 2], C.byref(tk), None) == 0
        pending.append(tk.value)
        if len(pending) == 3:
            assert sc.collect(pending.pop(0), copy=False).n_loci == slab
    while pending:
        assert sc.collect(pending.pop(0), copy=False).n_loci == slab
def run(count):
    th = [threading.Thread(target=reader, args=(t, count)) for t in range(T)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    return time.perf_counter() - t0
run(3)
for r in range(reps):
    dt = run(n_slabs)
    print(f"T={T} rep {r}: {dt / (n_slabs * T) * 1e3:.3f} ms per slab, {nb[0] * n_slabs * T / dt / 1e9:.2f} GB/s")
