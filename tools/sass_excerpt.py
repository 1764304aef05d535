#!/usr/bin/env python
"""SASS evidence without a GPU: for every kernel of interest in the built objects, the count of the mnemonics that
prove what the source claims (UBLKCP = cp.async.bulk, DMMA = FP64 tensor core, LDG.E.*.128 / LDS.128 = 128-bit loads,
SYNCS = mbarrier, BAR = named barriers) and the first lines that carry them.  usage: sass_excerpt.py > profiles/sass_rN.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "poolgen_b200", "lib", "obj")
WANT = [("pg_scan_a4.o", r"scan_kernelILi4ELi3ELb0ELi32ELi0"), ("pg_scan_a4.o", r"scan_kernelILi4ELi1ELb0ELi16ELi0"),
        ("pg_scan_a4.o", r"fixup_kernelILi4ELi3ELb0"), ("pg_kinship.o", r"gram_kernel"), ("pg_kinship.o", r"covar_mma_kernelILi2"),
        ("pg_kinship.o", r"covar_kernelILi2ELi1"), ("pg_tables.o", r"tables_kernel_tILi2ELi4"), ("pg_tables.o", r"fisher_wide_kernel"),
        ("pg_nm.o", r"mle_kernel"), ("pg_nm.o", r"gwalpha_kernel"), ("pg_text.o", r"text_parse_kernel"), ("pg_ingest.o", r"ingest_counts_kernelIh")]
KEYS = ["UBLKCP", "DMMA", "DFMA", "LDG.E", "LDS.128", "LDS.64", "STG.E", "SYNCS", "BAR", "MUFU", "SHFL", "ATOM", "RED"]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and "/*" in line and ";" in line:
            body[cur].append(line.strip())
    return body


def main():
    cache = {}
    for obj, pat in WANT:
        path = os.path.join(OBJ, obj)
        if path not in cache:
            cache[path] = functions(path)
        for name, lines in cache[path].items():
            if not re.search(pat, name):
                continue
            dem = subprocess.run(["cu++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip() or name
            print(f"== {dem[:150]}  ({obj}, {len(lines)} SASS instructions)")
            counts = {k: sum(1 for l in lines if re.search(r"\b" + re.escape(k), l)) for k in KEYS}
            print("   " + "  ".join(f"{k}={v}" for k, v in counts.items() if v))
            wide = sum(1 for l in lines if re.search(r"LDG\.E\S*\.128", l))
            if wide:
                print(f"   128-bit global loads: {wide}")
            for k in ("UBLKCP", "DMMA", "SYNCS", "BAR.SYNC", "BAR.ARV"):
                hits = [l for l in lines if k in l][:2]
                for h in hits:
                    print("      " + re.sub(r"\s+", " ", h)[:150])
            break
        else:
            print(f"== {pat}: not found in {obj}")


if __name__ == "__main__":
    main()
