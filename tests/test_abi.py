"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly what include/poolgen_cuda.h
declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import poolgen_b200 as pb
from poolgen_b200 import capi
from poolgen_b200.build import LIB_PATH, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(LIB_PATH):
        build()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "poolgen_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    declared = _header_functions()
    assert sorted(capi.ABI_SYMBOLS) == declared
    L = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert capi.lib().pg_abi_version() == 1


def test_synth_library_is_separate_and_cuda_free():
    """include/poolgen_synth.h lives in libpoolgen_synth.so: the host replay of the synthetic workload does not map the
    product library (bench.py's reference arm generates its inputs there)"""
    import subprocess
    from poolgen_b200.build import SYNTH_LIB_PATH
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "poolgen_synth.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(capi.SYNTH_SYMBOLS)
    L = ctypes.CDLL(SYNTH_LIB_PATH)
    P = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
        assert not hasattr(P, name), name
    out = subprocess.run(["ldd", SYNTH_LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "cuda" not in out.lower() and "poolgen" not in out, out


def test_shard_range_and_comm_id_without_a_gpu():
    """the host-callable part of the multi-GPU ABI: contiguous ranges and the communicator id (NCCL loads lazily)"""
    assert [pb.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert pb.shard_range(0, 0, 1) == (0, 0)
    with pytest.raises(pb.PgError):
        pb.shard_range(10, 3, 3)
    assert pb.nccl_version() >= 22000
    a, b = pb.Comm.unique_id(), pb.Comm.unique_id()
    assert len(a) == len(b) == capi.COMM_ID_BYTES and a != b


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pb.PgError) as e:
        pb.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poolgen_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU oracle", ""), os.path.join(dirpath, f)


def test_host_generator_is_deterministic_and_well_formed():
    a = pb.synth_counts_host(0x5EED0002, 10, 2000, 100, 4)
    b = pb.synth_counts_host(0x5EED0002, 10, 2000, 100, 4)
    assert np.array_equal(a, b)
    # shifting the window reproduces the same loci
    c = pb.synth_counts_host(0x5EED0002, 510, 100, 100, 4)
    assert np.array_equal(a[500:600], c)
    depth = a.sum(axis=1)
    assert depth.max() <= 100 and ((depth == 0) | (depth >= 20)).all()
    zero_loci = (depth == 0).any(axis=1).mean()
    mono = ((a > 0).any(axis=2).sum(axis=1) == 1).mean()
    assert 0.005 < zero_loci < 0.04 and 0.03 < mono < 0.08
    y = pb.synth_phen_host(1, 100, 3)
    assert y.shape == (100, 3) and (np.abs(y) <= 3).all() and y.std() > 1.0


def test_rust_ffi_crate_declares_the_same_functions():
    """ffi/poolgen-cuda-sys/src/lib.rs (the binding poolgen would link; not compiled here: no cargo in the image) must
    declare every function of the header with the same number of parameters, and nothing else"""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "poolgen_cuda.h")).read(), flags=re.S)
    rs = re.sub(r"//.*", "", open(os.path.join(ROOT, "ffi", "poolgen-cuda-sys", "src", "lib.rs")).read())
    ext = rs[rs.index('extern "C" {'):]
    ext = ext[:ext.index("\n}\n")]

    def arity(params: str) -> int:
        p = params.strip()
        return 0 if p in ("", "void") else p.count(",") + 1

    c_fns = {m.group(1): arity(m.group(2)) for m in re.finditer(r"\b(pg_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr)}
    rs_fns = {m.group(1): arity(m.group(2)) for m in re.finditer(r"pub fn (pg_[a-z0-9_]+)\(([^()]*)\)", ext)}
    assert sorted(c_fns) == _header_functions()
    assert rs_fns == c_fns


def test_library_does_not_link_the_cuda_math_libraries():
    """cuSOLVER is dlopen()ed by the eigen step only: a process that opens libpoolgen_cuda.so for the scans must not
    map cuSOLVER / cuSPARSE / cuBLAS (1.5 GB, minutes on a cold file system)"""
    import subprocess
    out = subprocess.run(["ldd", LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    for name in ("cusolver", "cublas", "cusparse", "nvJitLink", "nccl"):
        assert name not in out, out
