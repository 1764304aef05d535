"""File-level mirror of the reference's readers (poolgen_b200/sync_io.py): FilePhen::lparse, find_file_splits on the CPU,
FileSyncPhen / FileSync::read_analyse_write on the GPU."""
import json
import os

import numpy as np
import pytest

import poolgen_b200 as pb
from tests import helpers as H


def _write_c1_phen(path):
    with open(os.path.join(H.HERE, "golden", "c1_phen.json")) as fh:
        ph = json.load(fh)
    lines = ["#name,poolSizes,Trait_A,Trait_B"]
    for name, size, row in zip(ph["pool_names"], ph["pool_sizes_raw"], ph["phen"]):
        lines.append(f"{name}, {size:g},{row[0]} ,{row[1]}")
    with open(path, "w") as fh:
        fh.write("\r\n".join(lines) + "\r\n")


def test_file_phen_known_answer(tmp_path):
    # src/base/phen.rs:208-236: tests/test.csv -> names G1..G5, pool sizes 0.2 each, traits columns 2 and 3
    p = str(tmp_path / "test.csv")
    _write_c1_phen(p)
    phen = pb.FilePhen(p, ",", 0, 1, [2, 3]).lparse()
    assert phen.pool_names == ["G1", "G2", "G3", "G4", "G5"]
    assert list(phen.pool_sizes) == [0.2, 0.2, 0.2, 0.2, 0.2]
    assert phen.phen_matrix.T.tolist() == [[0.1, 0.3, 0.5, 0.7, 0.9], [83.2, 75.3, 49.8, 23.9, 12.0]]
    with open(p, "a") as fh:
        fh.write("G6,10,NA,\n")
    phen = pb.FilePhen(p, ",", 0, 1, [2, 3]).lparse()
    assert np.isnan(phen.phen_matrix[5]).all() and abs(phen.pool_sizes.sum() - 1.0) < 1e-15
    assert phen.pool_sizes[5] == 10.0 / 110.0


def test_find_file_splits(tmp_path):
    # src/base/helpers.rs:74-91: split points move to the start of the NEXT line (a full line is skipped even when the
    # raw split already is a line start), duplicates collapse
    p = str(tmp_path / "x.sync")
    lines = [("chr1\t%d\tN\t" % i) + "\t".join(["1:2:3:4:0:0"] * 3) + "\n" for i in range(100)]
    with open(p, "w") as fh:
        fh.writelines(lines)
    size = os.path.getsize(p)
    starts = np.cumsum([0] + [len(l) for l in lines])
    for n_threads in (1, 2, 3, 7):
        s = pb.find_file_splits(p, n_threads)
        assert s[0] == 0 and s[-1] == size and s == sorted(set(s))
        assert all(x in starts for x in s)
        raw = list(range(0, size, size // n_threads)) + [size]
        expect = [0 if r == 0 else int(starts[np.searchsorted(starts, r, side="right")]) if r < size else size for r in raw]
        dedup = [expect[0]] + [b for a, b in zip(expect, expect[1:]) if a != b]
        assert s == dedup
    with pytest.raises(pb.PgError):
        pb.find_file_splits(str(tmp_path / "missing.sync"), 2)


@pytest.mark.gpu
@pytest.mark.parametrize("analysis", ["ols_iter", "pearson_corr", "chisq_test", "fisher_exact_test"])
def test_read_analyse_write_c1(ctx, tmp_path, analysis):
    """`poolgen <analysis> -f tests/test.sync -p tests/test.csv --n-threads 3` through the mirror: three reader threads
    over line-aligned byte ranges, blocks of 48 KB (so lines straddle block ends), rows concatenated in chunk order;
    the file equals header + rows of the same records computed from the parsed counts in one batch"""
    from tests.test_text_gpu import _sync_text
    c1 = H.load_c1()
    names = [str(s) for s in c1["chrom_names"]]
    chroms = [names[i] for i in c1["chrom_idx"]]
    pos = [int(p) for p in c1["pos"]]
    fsync = str(tmp_path / "test.sync")
    with open(fsync, "wb") as fh:
        fh.write(_sync_text(c1["counts"], chroms, pos))
    fphen = str(tmp_path / "test.csv")
    _write_c1_phen(fphen)
    phen = pb.FilePhen(fphen, ",", 0, 1, [2, 3]).lparse()
    fs = pb.FilterStats(pool_sizes=phen.pool_sizes, min_coverage_depth=10, min_allele_frequency=0.01)
    out = str(tmp_path / "out.csv")
    if analysis in ("ols_iter", "pearson_corr"):
        fn = pb.ols_iterate if analysis == "ols_iter" else pb.correlation
        src = pb.FileSyncPhen(fsync, phen.pool_names, phen.pool_sizes, phen.phen_matrix, analysis)
        y = phen.phen_matrix
    else:
        fn = pb.chisq if analysis == "chisq_test" else pb.fisher
        src = pb.FileSync(fsync, analysis)
        y = None
    got_name = src.read_analyse_write(ctx, fs, out, 3, fn, block_bytes=48 << 10)
    assert got_name == out
    kind = {"ols_iter": pb.KIND_OLS, "pearson_corr": pb.KIND_CORR, "chisq_test": pb.KIND_CHISQ,
            "fisher_exact_test": pb.KIND_FISHER}[analysis]
    scan = pb.Scan(ctx, kind, fs, 5, c1["codes"], y)
    whole = scan.run_counts(c1["counts"])
    scan.close()
    expect = pb.format_header(kind) + pb.format_rows(kind, whole, c1["pos"], chr_names=names, chr_index=c1["chrom_idx"])
    got = open(out, "rb").read()
    assert got == expect and got.count(b"\n") > 1500
    with pytest.raises(pb.PgError):   # create_new(true): the output file must not exist
        src.read_analyse_write(ctx, fs, out, 3, fn)
    # default output name: <sync name without extension>-<seconds>-<test>.csv (src/base/sync.rs:885-903)
    auto = src.read_analyse_write(ctx, fs, "", 2, fn)
    assert auto.startswith(str(tmp_path / "test-")) and auto.endswith(f"-{analysis}.csv")
    assert open(auto, "rb").read() == expect
