"""File-level mirror of the reference's chunked readers for the hot path: `FilePhen::lparse` (src/base/phen.rs:21-98),
`find_file_splits` (src/base/helpers.rs:16-27, 74-91) and `ChunkyReadAnalyseWrite::read_analyse_write` of `FileSyncPhen`
/ `FileSync` (src/base/sync.rs:606-786, 788-970), with the same names and argument meaning.

What differs is where the work happens: a reader thread does not parse its lines and call a per-locus callback, it hands
line-aligned blocks of raw bytes to the library (`pg_scan_submit_sync_text`: parse + filter + regression on the GPU) and
appends the rows `pg_format_rows` returns.  Nothing here computes on the CPU; without the CUDA library every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import time
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import (KIND_CHISQ, KIND_CORR, KIND_FISHER, KIND_GWALPHA_LS, KIND_GWALPHA_ML, KIND_MLE, KIND_OLS, Context,
                   FilterStats, PgError, Scan)

_MISSING = {"", "NA", "NAN", "NaN", "na", "nan"}  # src/base/phen.rs:66-72


@dataclass
class Phen:
    """src/base/structs_and_traits.rs: Phen { pool_names, pool_sizes (normalised to sum 1), phen_matrix n x k }"""
    pool_names: list
    pool_sizes: np.ndarray
    phen_matrix: np.ndarray


@dataclass
class FilePhen:
    """src/base/structs_and_traits.rs: FilePhen; only the "default" delimited format (the GWAlpha format feeds
    `gwalpha`, which is not part of this path)"""
    filename: str
    delim: str = ","
    names_column_id: int = 0
    sizes_column_id: int = 1
    trait_values_column_ids: list = field(default_factory=lambda: [2])
    format: str = "default"

    def lparse(self) -> Phen:
        """src/base/phen.rs:21-98: '#' lines skipped, fields trimmed, missing values -> NaN, pool sizes divided by their
        sequential sum"""
        if self.format != "default":
            raise PgError("Invalid phenotype format for this path: only 'default' (src/base/phen.rs:24-98)")
        names, sizes, vals = [], [], []
        with open(self.filename, "r", newline="") as fh:
            for line in fh:
                line = line.rstrip("\n").rstrip("\r")
                if line[:1] == "#":  # the reference indexes byte 0: an empty line panics there
                    continue
                if line == "":
                    raise PgError("empty line in the phenotype file (the reference panics: index out of bounds)")
                f = [x.strip() for x in line.split(self.delim)]
                names.append(f[self.names_column_id])
                try:
                    sizes.append(float(f[self.sizes_column_id]))
                except ValueError:
                    raise PgError(f"T_T Pool sizes column (column index: {self.sizes_column_id}) is not a valid number. "
                                  f"Line: {line}.") from None
                for j in self.trait_values_column_ids:
                    vals.append(float("nan") if f[j] in _MISSING else float(f[j]))
        total = 0.0
        for s in sizes:  # `pool_sizes.iter().sum()`: sequential
            total = total + s
        k = len(self.trait_values_column_ids)
        n = len(vals) // k
        return Phen(names, np.array([s / total for s in sizes], dtype=np.float64),
                    np.array(vals, dtype=np.float64).reshape(n, k))


def _find_start_of_next_line(fname: str, pos: int) -> int:
    """src/base/helpers.rs:16-27: for pos > 0 the rest of the line at pos is skipped (a full line when pos is a start)"""
    if pos <= 0:
        return 0
    with open(fname, "rb") as fh:
        fh.seek(pos)
        fh.readline()
        return fh.tell()


def find_file_splits(fname: str, n_threads: int) -> list:
    """src/base/helpers.rs:74-91: (0..end).step_by(end / n_threads) + end, each moved to the start of the next line,
    consecutive duplicates removed"""
    if not os.path.exists(fname):
        raise PgError(f"The input file: {fname} does not exist. Please make sure you are entering the correct filename "
                      "and/or the correct path.")
    end = os.path.getsize(fname)
    step = end // n_threads
    if step == 0:
        raise PgError("more threads than bytes in the input file (the reference panics: step_by(0))")
    out = list(range(0, end, step)) + [end]
    out = [_find_start_of_next_line(fname, p) for p in out]
    dedup = [out[0]]
    for p in out[1:]:
        if p != dedup[-1]:
            dedup.append(p)
    return dedup


_KIND_NAMES = {KIND_OLS: "ols_iter", KIND_CORR: "pearson_corr", KIND_CHISQ: "chisq_test", KIND_FISHER: "fisher_exact_test",
               KIND_MLE: "mle_iter", KIND_GWALPHA_LS: "gwalpha", KIND_GWALPHA_ML: "gwalpha"}


def _kind_of(function) -> int:
    """the reference passes the callback itself (gwas::ols_iterate, gwas::correlation, tables::chisq, tables::fisher)"""
    if isinstance(function, int):
        return function
    name = getattr(function, "__name__", str(function))
    table = {"ols_iterate": KIND_OLS, "correlation": KIND_CORR, "chisq": KIND_CHISQ, "fisher": KIND_FISHER,
             "mle_iterate": KIND_MLE, "gwalpha_ls": KIND_GWALPHA_LS, "gwalpha_ml": KIND_GWALPHA_ML}
    if name not in table:
        raise PgError(f"unknown per-locus function {name!r}")
    return table[name]


def _default_out(fname: str, test: str) -> str:
    """src/base/sync.rs:885-903: <name without its last extension>-<seconds>-<test>.csv"""
    bname = ".".join(fname.split(".")[:-1])
    return f"{bname}-{time.time()}-{test}.csv"


def _chunk_worker(ctx: Context, kind: int, fs: FilterStats, n_pools: int, phen, fname: str, start: int, end: int,
                  block_bytes: int, fmt_threads: int, sink: list, errors: list):
    """the body of `per_chunk` (src/base/sync.rs:794-870 / 612-687) for the byte range [start, end)"""
    lib = capi.lib()
    scan = None
    pinned = []
    try:
        scan = Scan(ctx, kind, fs, n_pools, np.arange(6, dtype=np.uint8), phen)
        # a locus line holds at least "c\\t0\\tN" + n_pools * "\\t0:0:0:0:0:0" bytes
        max_loci = block_bytes // (12 * n_pools + 5) + 2
        scan.stream_begin(max_loci)
        bufs = []
        for _ in range(capi_stream_depth()):
            arr, h = ctx.pinned_empty((block_bytes,), np.uint8)
            pinned.append(h)
            bufs.append(arr)
        out = bytearray()
        pending = []  # (ticket, buffer index)

        def finish(item):
            ticket, bi = item
            res = scan.collect(ticket, copy=False)
            if res.n_loci == 0:
                return
            po, pp = C.c_void_p(), C.c_void_p()
            capi._check(lib.pg_scan_text_labels(scan._h, int(ticket), C.byref(po), C.byref(pp)), ctx._h,
                        "pg_scan_text_labels")
            lab = capi._RowLabels()
            lab.positions = C.cast(pp, C.POINTER(C.c_uint64))
            lab.text = C.cast(bufs[bi].ctypes.data, C.c_char_p)
            lab.line_offsets = C.cast(po, C.POINTER(C.c_uint64))
            need = C.c_size_t()
            # the output file carries the reference's own p-value digits (PG_FORMAT_EXACT_P)
            lib.pg_format_rows_ex(kind, C.byref(res), C.byref(lab), capi.FORMAT_EXACT_P, scan.n_pools, fmt_threads,
                                  None, 0, C.byref(need))
            if need.value:
                rows = C.create_string_buffer(need.value)
                capi._check(lib.pg_format_rows_ex(kind, C.byref(res), C.byref(lab), capi.FORMAT_EXACT_P, scan.n_pools,
                                                  fmt_threads, rows, need.value, C.byref(need)), ctx._h,
                            "pg_format_rows_ex")
                out.extend(rows.raw[:need.value])

        with open(fname, "rb", buffering=0) as fh:
            fh.seek(start)
            left = end - start
            carry = b""
            bi = 0
            while left > 0 or carry:
                buf = bufs[bi]
                nc = len(carry)
                if nc:
                    buf[:nc] = np.frombuffer(carry, dtype=np.uint8)
                want = min(left, block_bytes - nc)
                got = fh.readinto(memoryview(buf)[nc:nc + want]) if want > 0 else 0
                left -= got
                fill = nc + got
                if left > 0 and got > 0:
                    # cut at the last newline; the tail opens the next block
                    view = buf[:fill]
                    nl = np.flatnonzero(view[::-1] == 10)
                    if nl.size == 0:
                        raise PgError(f"a line of {fname} is longer than the block size {block_bytes}")
                    cut = fill - int(nl[0])
                    carry = bytes(view[cut:])
                    fill = cut
                else:
                    carry = b""
                    if got == 0 and left > 0:
                        left = 0  # the file shrank
                ticket = C.c_int()
                capi._check(lib.pg_scan_submit_sync_text(scan._h, buf.ctypes.data, fill, C.byref(ticket), None), ctx._h,
                            "pg_scan_submit_sync_text")
                pending.append((ticket.value, bi))
                bi = (bi + 1) % len(bufs)
                if len(pending) == len(bufs):
                    finish(pending.pop(0))
            while pending:
                finish(pending.pop(0))
        sink.append((start, bytes(out)))
    except Exception as e:  # noqa: BLE001 -- re-raised by the caller, like a panicking reader thread
        errors.append(e)
    finally:
        if scan is not None:
            scan.close()
        for h in pinned:
            ctx.pinned_free(h)


def capi_stream_depth() -> int:
    return 3  # PG_STREAM_DEPTH


def _read_analyse_write(ctx, kind, fs, n_pools, phen, fname, test, out, n_threads, block_bytes):
    if out == "":
        out = _default_out(fname, test)
    if os.path.exists(out):
        raise PgError("Cannot write to output file")  # create_new(true), src/base/sync.rs:905
    chunks = find_file_splits(fname, n_threads)
    if len(chunks) - 1 < n_threads:
        # the reference indexes chunks[n_threads] unguarded (src/base/sync.rs:909) and panics
        raise PgError(f"{n_threads} threads but only {len(chunks) - 1} line-aligned chunks in {fname}")
    sink, errors, threads = [], [], []
    fmt_threads = max(1, (os.cpu_count() or 1) // n_threads)
    for i in range(n_threads):
        t = threading.Thread(target=_chunk_worker, args=(ctx, kind, fs, n_pools, phen, fname, chunks[i], chunks[i + 1],
                                                         block_bytes, fmt_threads, sink, errors))
        t.start()
        threads.append(t)
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    with open(out, "xb") as fo:
        fo.write(capi.format_header(kind))
        for _, rows in sorted(sink):  # chunk files are concatenated in name order = start offset order (sync.rs:953-967)
            fo.write(rows)
    return out


@dataclass
class FileSyncPhen:
    """src/base/structs_and_traits.rs: FileSyncPhen { filename_sync, pool_names, pool_sizes, phen_matrix, test }"""
    filename_sync: str
    pool_names: list
    pool_sizes: np.ndarray
    phen_matrix: np.ndarray
    test: str = "ols_iter"

    def read_analyse_write(self, ctx: Context, filter_stats: FilterStats, out: str, n_threads: int, function,
                           block_bytes: int = 32 << 20) -> str:
        """`read_analyse_write(&self, &FilterStats, out, n_threads, function)` (src/base/sync.rs:872-970) for
        function = ols_iterate | correlation | mle_iterate | gwalpha_ls | gwalpha_ml (for the last two phen_matrix is
        the gwalpha_fmt matrix, src/main.rs:335-358): returns the output file name"""
        kind = _kind_of(function)
        if kind not in (KIND_OLS, KIND_CORR, KIND_MLE, KIND_GWALPHA_LS, KIND_GWALPHA_ML):
            raise PgError("FileSyncPhen::read_analyse_write takes gwas::ols_iterate, correlation, mle_iterate or gwalpha_ls / _ml")
        return _read_analyse_write(ctx, kind, filter_stats, len(self.pool_names), self.phen_matrix, self.filename_sync,
                                   self.test, out, n_threads, block_bytes)


@dataclass
class FileSync:
    """src/base/structs_and_traits.rs: FileSync { filename, test } -- the count tests need no phenotypes, only the pool
    sizes that travel in FilterStats"""
    filename: str
    test: str = "chisq_test"

    def read_analyse_write(self, ctx: Context, filter_stats: FilterStats, out: str, n_threads: int, function,
                           block_bytes: int = 32 << 20) -> str:
        """src/base/sync.rs:689-786 for function = chisq | fisher"""
        kind = _kind_of(function)
        if kind not in (KIND_CHISQ, KIND_FISHER):
            raise PgError("FileSync::read_analyse_write takes tables::chisq or tables::fisher")
        return _read_analyse_write(ctx, kind, filter_stats, int(np.asarray(filter_stats.pool_sizes).size), None,
                                   self.filename, self.test, out, n_threads, block_bytes)


# ---- whole-matrix entries: sync2csv (SaveCsv::write_csv, src/base/sync.rs:1182-1262) and ols_iter_with_kinship
# (into_genotypes_and_phenotypes + ols_with_covariate, src/base/sync.rs:1106-1179, src/gwas/ols.rs:278-436) ----------
@dataclass
class _LoadedColumns:
    kin: object            # capi.Kinship holding the allele columns on the device, in file order
    col_locus: np.ndarray  # int64 [P] locus ordinal (over the kept loci of the whole file) of every column
    col_allele: np.ndarray # uint8 [P]
    chr_names: list        # distinct chromosome names
    chr_index: np.ndarray  # uint32 [L]
    positions: np.ndarray  # uint64 [L]


def _load_all(ctx: Context, fname: str, n_pools: int, fs: FilterStats, keep_p_minus_1: bool, max_columns: int,
              block_bytes: int) -> _LoadedColumns:
    """LoadAll::load over the whole file, block by block through the device parser and loader"""
    kin = capi.Kinship(ctx, n_pools, max_columns)
    max_loci = block_bytes // (12 * n_pools + 5) + 2
    names, index_of = [], {}
    chr_index, positions, col_locus, col_allele = [], [], [], []
    base = 0
    try:
        with open(fname, "rb") as fh:
            carry = b""
            while True:
                data = fh.read(block_bytes)
                if not data and not carry:
                    break
                block = carry + data
                if data:
                    cut = block.rfind(b"\n") + 1
                    if cut == 0:
                        raise PgError(f"a line of {fname} is longer than the block size {block_bytes}")
                    block, carry = block[:cut], block[cut:]
                else:
                    carry = b""
                L, off, pos, loc, alle = kin.append_sync_text(block, fs, max_loci, keep_p_minus_1)
                for o in off:  # chromosome = the text up to the first tab of the locus' line
                    o = int(o)
                    nm = block[o:block.index(b"\t", o)]
                    if nm not in index_of:
                        index_of[nm] = len(names)
                        names.append(nm)
                    chr_index.append(index_of[nm])
                positions.append(pos)
                col_locus.append(loc + base)
                col_allele.append(alle)
                base += L
                if not data:
                    break
    except Exception:
        kin.close()
        raise
    cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt)
    return _LoadedColumns(kin, cat(col_locus, np.int64), cat(col_allele, np.uint8), names,
                          np.array(chr_index, dtype=np.uint32), cat(positions, np.uint64))


def _count_columns_bound(fname: str) -> int:
    """an upper bound of the allele columns LoadAll can emit: five per line"""
    with open(fname, "rb") as fh:
        lines = sum(chunk.count(b"\n") for chunk in iter(lambda: fh.read(1 << 24), b"")) + 1
    return 5 * lines


def write_csv(self, ctx: Context, filter_stats: FilterStats, keep_p_minus_1: bool, out: str, n_threads: int,
              block_bytes: int = 32 << 20) -> str:
    """sync2csv: `SaveCsv::write_csv(&self, &FilterStats, keep_p_minus_1, out, n_threads)` of FileSyncPhen
    (src/base/sync.rs:1182-1262): header `#chr,pos,allele,<pool names>`, one row per kept allele of every kept locus,
    loci in LoadAll's order (stable by chromosome, position), frequencies through parse_f64_roundup_and_own(x, 6)"""
    if out == "":
        bname = ".".join(self.filename_sync.split(".")[:-1])
        out = f"{bname}-{time.time()}-allele_frequencies.csv"
    if os.path.exists(out):
        raise PgError("Cannot write to output file")
    ld = _load_all(ctx, self.filename_sync, len(self.pool_names), filter_stats, keep_p_minus_1,
                   _count_columns_bound(self.filename_sync), block_bytes)
    try:
        if ld.col_locus.size == 0:
            raise PgError("No data passed the filtering variables. Please decrease minimum depth, and/or minimum "
                          "allele frequency.")  # assert at src/base/sync.rs:1219
        G = ld.kin.get_columns(0, ld.kin.columns)
    finally:
        ld.kin.close()
    order = capi.sort_loci(ld.positions, chr_names=ld.chr_names, chr_index=ld.chr_index)
    rows = capi.format_frequency_rows(G, ld.col_locus, ld.col_allele, ld.positions, locus_order=order,
                                      n_threads=max(1, n_threads), chr_names=ld.chr_names, chr_index=ld.chr_index)
    with open(out, "xb") as fo:
        fo.write(capi.format_frequency_header(self.pool_names))
        fo.write(rows)
    return out


def _iter_with_kinship(self, ctx: Context, filter_stats: FilterStats, keep_p_minus_1: bool,
                       xxt_eigen_variance_explained: float, out: str, n_threads: int, block_bytes: int, method: str) -> str:
    if out != "" and os.path.exists(out):
        raise PgError("Cannot write to output file")
    if np.isnan(self.phen_matrix).any():
        raise PgError("pools with missing phenotypes: remove them before the scan (the reference's remove_missing, "
                      "src/gwas/ols.rs:286)")
    ld = _load_all(ctx, self.filename_sync, len(self.pool_names), filter_stats, keep_p_minus_1,
                   _count_columns_bound(self.filename_sync), block_bytes)
    try:
        P = ld.kin.columns
        if P == 0:
            raise PgError("No data passed the filtering variables.")
        ld.kin.gram()
        n_eigenvecs = ld.kin.eig_select(P, float(xxt_eigen_variance_explained))
        scan = ld.kin.covar_scan if method == "ols" else ld.kin.mle_scan
        beta, _var, pval = scan(self.phen_matrix)   # [k, P] in file order
    finally:
        ld.kin.close()
    # the matrix columns follow the loci sorted by (chromosome, position): permute the records into that order
    order = capi.sort_loci(ld.positions, chr_names=ld.chr_names, chr_index=ld.chr_index)
    first = np.full(int(ld.positions.size) + 1, -1, dtype=np.int64)
    count = np.zeros(int(ld.positions.size) + 1, dtype=np.int64)
    for c, l in enumerate(ld.col_locus):
        if first[l] < 0:
            first[l] = c
        count[l] += 1
    seq = np.concatenate([np.arange(first[l], first[l] + count[l]) for l in order if first[l] >= 0]).astype(np.int64)
    names = [n.decode() for n in ld.chr_names]
    chromosome = ["intercept"] + [names[ld.chr_index[ld.col_locus[c]]] for c in seq]
    position = [0] + [int(ld.positions[ld.col_locus[c]]) for c in seq]
    allele = ["intercept"] + [capi.ALLELE_NAMES[ld.col_allele[c]] for c in seq]
    rows = capi.format_kinship_rows(chromosome, position, allele, beta[:, seq], pval[:, seq], n_threads=max(1, n_threads))
    if out == "":  # src/gwas/ols.rs:374-398, src/gwas/mle.rs:409-433
        bname = ".".join(self.filename_sync.split(".")[:-1])
        out = f"{bname}-{method}_iterative_xxt_{n_eigenvecs + 1}_eigens-{time.time()}.csv"
    with open(out, "xb") as fo:
        fo.write(capi.format_header(capi.KIND_OLS_KINSHIP))
        fo.write(rows)
    return out


def ols_iter_with_kinship(self, ctx: Context, filter_stats: FilterStats, keep_p_minus_1: bool,
                          xxt_eigen_variance_explained: float, out: str, n_threads: int,
                          block_bytes: int = 32 << 20) -> str:
    """`poolgen ols_iter_with_kinship` (src/main.rs:280-298): into_genotypes_and_phenotypes (src/base/sync.rs:1106-1179)
    then ols_with_covariate (src/gwas/ols.rs:278-436) and its writer, including the label indexing of
    src/gwas/ols.rs:421-424 (row i of the allele columns carries entry i of the label vectors, whose entry 0 is the
    intercept's)"""
    return _iter_with_kinship(self, ctx, filter_stats, keep_p_minus_1, xxt_eigen_variance_explained, out, n_threads,
                              block_bytes, "ols")


def mle_iter_with_kinship(self, ctx: Context, filter_stats: FilterStats, keep_p_minus_1: bool,
                          xxt_eigen_variance_explained: float, out: str, n_threads: int,
                          block_bytes: int = 32 << 20) -> str:
    """`poolgen mle_iter_with_kinship` (src/main.rs:316-326): the same loader, then mle_with_covariate
    (src/gwas/mle.rs:307-463) -- kinship matrix, PCs by the same rule, one maximum-likelihood fit per column and
    phenotype -- and its writer (same header and row layout, labels indexed the same way, mle.rs:449-460)"""
    return _iter_with_kinship(self, ctx, filter_stats, keep_p_minus_1, xxt_eigen_variance_explained, out, n_threads,
                              block_bytes, "mle")


FileSyncPhen.write_csv = write_csv
FileSyncPhen.ols_iter_with_kinship = ols_iter_with_kinship
FileSyncPhen.mle_iter_with_kinship = mle_iter_with_kinship
