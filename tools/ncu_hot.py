#!/usr/bin/env python
"""Summarises an ncu report's SASS page without a GPU: per address range (cut at calls/returns or fixed buckets) the
share of executed warp instructions and of stall samples, and the top stalled instructions with their reasons.

usage: ncu_hot.py report.ncu-rep [n_loci] [kernel_index] [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    loci = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    s, e = starts[kidx], starts[kidx + 1]
    print(rows[s][1][:100])
    h = rows[s + 1]
    ia, isrc, ie, isamp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for r in rows[s + 2:e]:
        try:
            data.append((int(r[ia], 16), r[isrc], int(r[ie]), int(r[isamp]), r))
        except (ValueError, IndexError):
            pass
    base = data[0][0]
    tot = sum(d[2] for d in data)
    tots = sum(d[3] for d in data)
    print(f"warp instructions {tot} ({tot / loci:.1f} per locus), samples {tots}")
    # function boundaries: an instruction after RET/EXIT/BRA-to-self starts a new range
    ranges, cur = [], [data[0][0] - base, 0, 0, 0]
    for a, src, ex, sm, _ in data:
        cur[1] += ex
        cur[2] += sm
        cur[3] = a - base
        if src.startswith("RET") or src.startswith("EXIT"):
            ranges.append(tuple(cur))
            cur = [a - base + 16, 0, 0, 0]
    ranges.append(tuple(cur))
    print("ranges (cut at RET/EXIT): start-end  inst%  samp%  inst/locus")
    for a0, ex, sm, a1 in ranges:
        if ex / max(tot, 1) > 0.002 or sm / max(tots, 1) > 0.002:
            print(f"  {a0:#8x}-{a1:#8x} {100 * ex / tot:6.2f} {100 * sm / tots:6.2f} {ex / loci:9.1f}")
    agg = {}
    for i, c in stall_cols:
        agg[c] = sum(int(d[4][i] or 0) for d in data)
    print("stall mix:", ", ".join(f"{c[6:]} {100 * v / max(tots, 1):.1f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print("top stalled instructions:")
    for d in sorted(sorted(data, key=lambda d: -d[3])[:top_n], key=lambda d: d[0]):
        st = sorted(((int(d[4][i] or 0), c) for i, c in stall_cols), reverse=True)[:2]
        print(f"  {d[0] - base:#8x} {100 * d[3] / tots:5.2f}% ex/locus={d[2] / loci:7.3f} {d[1][:58]:58s} "
              f"{[(c[6:], v) for v, c in st if v]}")


if __name__ == "__main__":
    main()
