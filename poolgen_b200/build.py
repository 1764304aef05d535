"""Builds libpoolgen_cuda.so (nvcc, sm_100a) in-tree: poolgen_b200/lib/libpoolgen_cuda.so."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "lib", "libpoolgen_cuda.so")
SYNTH_LIB_PATH = os.path.join(_HERE, "lib", "libpoolgen_synth.so")  # host replay of the synthetic workload (no CUDA)


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(SYNTH_LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs.append(os.path.join(_HERE, "..", "include", "poolgen_cuda.h"))
    srcs.append(os.path.join(_HERE, "..", "include", "poolgen_synth.h"))
    return any(os.path.getmtime(s) > t for s in srcs if os.path.isfile(s))


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a.  Raises if nvcc is missing or the build fails."""
    if force or _stale():
        jobs = jobs or max(1, (os.cpu_count() or 2))
        cmd = ["make", "-C", CSRC, f"-j{jobs}"] + (["-B"] if force else [])
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout)
        if res.returncode != 0:
            raise RuntimeError("building libpoolgen_cuda.so failed (see output above)")
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True))
