#!/usr/bin/env python
"""Concurrent pinned host-to-device copies on 1..N GPUs, one process per GPU: the host-side ceiling behind the
end-to-end leg of bench.py at N > 1.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port P tools/h2d_probe.py [--mb 1024] [--seconds 2]; rank 0 prints one JSON line per
phase: every rank copying at once, then each rank alone in turn."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def copy_rate(dev_buf, host_buf, seconds):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t_end = time.perf_counter() + seconds
    e0.record()
    while time.perf_counter() < t_end:
        for _ in range(4):
            dev_buf.copy_(host_buf, non_blocking=True)
            n += 1
        torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    return n * host_buf.numel() * host_buf.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=2.0)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    host = torch.empty(a.mb << 20, dtype=torch.uint8).pin_memory()
    host.fill_(7)
    dev = torch.empty(a.mb << 20, dtype=torch.uint8, device="cuda")
    copy_rate(dev, host, 0.3)
    if world > 1:
        dist.barrier()
    together = copy_rate(dev, host, a.seconds)
    rates = [None] * world
    if world > 1:
        dist.all_gather_object(rates, together)
    else:
        rates = [together]
    alone = []
    for r in range(world):
        if world > 1:
            dist.barrier()
        v = copy_rate(dev, host, min(a.seconds, 1.0)) if r == rank else 0.0
        if world > 1:
            box = [None] * world
            dist.all_gather_object(box, v)
            v = box[r]
        alone.append(v)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "buffer_mb": a.mb, "vcpus": os.cpu_count(),
                          "concurrent_gb_per_s_per_gpu": [round(x, 2) for x in rates],
                          "concurrent_gb_per_s_total": round(sum(rates), 2),
                          "alone_gb_per_s_per_gpu": [round(x, 2) for x in alone]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
