"""Regenerates the committed fixtures under tests/golden/ from the reference's own test data.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs
    c1_sync.npz   parsed form of /root/reference/tests/test.sync (config C1: 5 pools x 6674 loci):
                  chrom_names, chrom_idx[L], pos[L], counts[L, n_pools, 6] (sync column order A:T:C:G:N:D)
    c1_phen.json  /root/reference/tests/test.csv columns used by the reference's CI runs
                  (--phen-name-col 0 --phen-pool-size-col 1 --phen-value-col 2,3)
The parse here is a plain split on tabs/colons and is cross-checked against the oracle's
restatement of sync.rs:100-156 for every line before anything is written.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pgo  # noqa: E402

REF = "/root/reference/tests"


def main():
    chrom_names, chrom_idx, pos, counts = [], [], [], []
    with open(os.path.join(REF, "test.sync")) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            f = line.rstrip("\n").rstrip("\r").split("\t")
            c = np.array([[int(v) for v in pool.split(":")[:6]] for pool in f[3:]], dtype=np.uint64)
            n, ch, p, oc = pgo.parse_sync_line(line)
            assert n == c.shape[0] and ch == f[0] and p == int(f[1]) and (oc == c).all()
            if f[0] not in chrom_names:
                chrom_names.append(f[0])
            chrom_idx.append(chrom_names.index(f[0]))
            pos.append(int(f[1]))
            counts.append(c)
    counts = np.stack(counts)
    assert counts.max() < 65536
    np.savez_compressed(os.path.join(HERE, "c1_sync.npz"), chrom_names=np.array(chrom_names),
                        chrom_idx=np.array(chrom_idx, dtype=np.int32), pos=np.array(pos, dtype=np.uint64),
                        counts=counts.astype(np.uint16))
    names, sizes, traits = [], [], []
    with open(os.path.join(REF, "test.csv")) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            f = [x.strip() for x in line.rstrip("\n").split(",")]
            names.append(f[0]); sizes.append(float(f[1])); traits.append([float(f[2]), float(f[3])])
    with open(os.path.join(HERE, "c1_phen.json"), "w") as fh:
        json.dump({"pool_names": names, "pool_sizes_raw": sizes, "phen": traits,
                   "source": "poolgen tests/test.csv, columns 0,1,2,3"}, fh, indent=1)
    print("loci", counts.shape, "chromosomes", len(chrom_names))


if __name__ == "__main__":
    main()
