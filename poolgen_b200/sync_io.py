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
from .capi import KIND_CHISQ, KIND_CORR, KIND_FISHER, KIND_OLS, Context, FilterStats, PgError, Scan

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


_KIND_NAMES = {KIND_OLS: "ols_iter", KIND_CORR: "pearson_corr", KIND_CHISQ: "chisq_test", KIND_FISHER: "fisher_exact_test"}


def _kind_of(function) -> int:
    """the reference passes the callback itself (gwas::ols_iterate, gwas::correlation, tables::chisq, tables::fisher)"""
    if isinstance(function, int):
        return function
    name = getattr(function, "__name__", str(function))
    table = {"ols_iterate": KIND_OLS, "correlation": KIND_CORR, "chisq": KIND_CHISQ, "fisher": KIND_FISHER}
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
            lib.pg_format_rows(kind, C.byref(res), C.byref(lab), fmt_threads, None, 0, C.byref(need))
            if need.value:
                rows = C.create_string_buffer(need.value)
                capi._check(lib.pg_format_rows(kind, C.byref(res), C.byref(lab), fmt_threads, rows, need.value,
                                               C.byref(need)), ctx._h, "pg_format_rows")
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
        function = ols_iterate | correlation: returns the output file name"""
        kind = _kind_of(function)
        if kind not in (KIND_OLS, KIND_CORR):
            raise PgError("FileSyncPhen::read_analyse_write takes gwas::ols_iterate or gwas::correlation")
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
