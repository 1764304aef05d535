#!/usr/bin/env python
"""Joins an ncu SASS-page CSV (`ncu -i X.ncu-rep --page source --csv --print-source sass`) with `nvdisasm -g`
line info of the same kernel and prints executed instructions / stall samples per source line.

usage: ncu_lines.py sass.csv object.o 'mangled_kernel_name' [top_n]
"""
import csv
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict


def line_table(obj, mangled):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", cub], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    tab, cur, on = [], None, False
    for ln in txt:
        if ln.startswith(".text."):
            on = ln.strip() == f".text.{mangled}:"
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            tab.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return tab


def main():
    sass_csv, obj, mangled = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    h = rows[1]
    ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ia], 16), int(r[ie]), int(r[isamp])))
        except (ValueError, IndexError):
            pass
    base = data[0][0]
    tab = {off: (cur, txt) for off, cur, txt in line_table(obj, mangled)}
    per = defaultdict(lambda: [0, 0])
    tot_e = tot_s = 0
    miss = 0
    for a, e, s in data:
        cur = tab.get(a - base)
        key = cur[0] if cur else None
        if cur is None:
            miss += e
        per[key][0] += e
        per[key][1] += s
        tot_e += e
        tot_s += s
    print(f"total warp-instructions {tot_e}, samples {tot_s}, unmatched instr {miss}")
    print(f"{'line':>22} {'inst%':>7} {'samp%':>7}   instructions")
    for key, (e, s) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{str(key):>22} {100 * e / tot_e:7.2f} {100 * s / max(tot_s, 1):7.2f}   {e}")


if __name__ == "__main__":
    main()
