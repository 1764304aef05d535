"""poolgen_b200 -- B200 (sm_100a) implementation of poolgen's per-locus GWAS scan.

Host-side mirror of the reference's callback interface for the hot path only:
`ols_iterate` (src/gwas/ols.rs:201-276), `correlation` (src/gwas/correlation_test.rs:73-129),
`chisq` (src/tables/chisq_test.rs:5-47), `fisher` (src/tables/fisher_exact_test.rs:32-130), each taking a
BATCH of parsed loci instead of one.  All arithmetic runs in libpoolgen_cuda.so (hand-written CUDA);
nothing here computes on the CPU.
"""
from . import capi  # noqa: F401
from .capi import (ALLELE_NAMES, KIND_CHISQ, KIND_CORR, KIND_FISHER, KIND_GWALPHA_LS, KIND_GWALPHA_ML, KIND_MLE, KIND_OLS, LOCUS_FAILED, LOCUS_FILTERED,
                   LOCUS_OK, LOCUS_PANIC, LOCUS_UNSUPPORTED, Batch, Comm, Context, FilterStats, Kinship, PgError, Scan,
                   ScanResults, format_f64, format_frequency_header, format_frequency_rows, format_header, format_kinship_rows,
                   format_rows, nccl_version, shard_range, sort_loci, synth_counts_host, synth_phen_host, synth_sync_text_host)

from .sync_io import FilePhen, FileSync, FileSyncPhen, Phen, find_file_splits  # noqa: E402

__all__ = ["FilePhen", "FileSync", "FileSyncPhen", "Phen", "find_file_splits", "ALLELE_NAMES", "KIND_CHISQ", "KIND_CORR", "KIND_FISHER", "KIND_OLS", "LOCUS_FAILED", "LOCUS_FILTERED",
           "LOCUS_OK", "LOCUS_PANIC", "LOCUS_UNSUPPORTED", "Batch", "Context", "FilterStats", "Kinship", "PgError", "Scan",
           "ScanResults", "synth_counts_host", "synth_phen_host", "synth_sync_text_host", "ols_iterate", "correlation", "chisq", "fisher",
           "ols_with_covariate", "format_f64", "format_header", "format_kinship_rows", "format_rows", "format_frequency_header",
           "format_frequency_rows", "sort_loci", "Comm", "shard_range", "nccl_version", "KIND_MLE", "KIND_GWALPHA_LS",
           "KIND_GWALPHA_ML", "mle_iterate", "gwalpha", "gwalpha_ls", "gwalpha_ml"]

_SYNC_CODES = (0, 1, 2, 3, 4, 5)


def _run(kind, ctx, counts, filter_stats, phen, allele_codes):
    scan = Scan(ctx, kind, filter_stats, counts.shape[2], allele_codes, phen)
    try:
        return scan.run_counts(counts)
    finally:
        scan.close()


def ols_iterate(ctx, counts, phen, filter_stats, allele_codes=_SYNC_CODES):
    """gwas::ols_iterate over a batch: counts uint32 [L, A, n_pools], phen [n_pools, k]."""
    return _run(KIND_OLS, ctx, counts, filter_stats, phen, allele_codes)


def correlation(ctx, counts, phen, filter_stats, allele_codes=_SYNC_CODES):
    """gwas::correlation over a batch."""
    return _run(KIND_CORR, ctx, counts, filter_stats, phen, allele_codes)


def chisq(ctx, counts, filter_stats, allele_codes=_SYNC_CODES):
    """tables::chisq over a batch."""
    return _run(KIND_CHISQ, ctx, counts, filter_stats, None, allele_codes)


def fisher(ctx, counts, filter_stats, allele_codes=_SYNC_CODES):
    """tables::fisher over a batch."""
    return _run(KIND_FISHER, ctx, counts, filter_stats, None, allele_codes)


def mle_iterate(ctx, counts, phen, filter_stats, allele_codes=_SYNC_CODES):
    """gwas::mle_iterate over a batch (src/gwas/mle.rs:232-305): stats[..., 0] = beta, 1 = v_b, 2 = beta / v_b, 3 = p."""
    return _run(KIND_MLE, ctx, counts, filter_stats, phen, allele_codes)


def gwalpha(ctx, counts, gwalpha_fmt, filter_stats, method="LS", allele_codes=_SYNC_CODES):
    """gwas::gwalpha_ls / gwalpha_ml over a batch (src/gwas/gwalpha.rs:282-386); gwalpha_fmt [rows, 3]: column 0 bins,
    column 1 q, column 2 = sig, MIN, MAX, then -inf.  stats[..., 0, 0] = alpha."""
    return _run(KIND_GWALPHA_LS if method == "LS" else KIND_GWALPHA_ML, ctx, counts, filter_stats, gwalpha_fmt, allele_codes)


def gwalpha_ls(ctx, counts, gwalpha_fmt, filter_stats, allele_codes=_SYNC_CODES):
    return gwalpha(ctx, counts, gwalpha_fmt, filter_stats, "LS", allele_codes)


def gwalpha_ml(ctx, counts, gwalpha_fmt, filter_stats, allele_codes=_SYNC_CODES):
    return gwalpha(ctx, counts, gwalpha_fmt, filter_stats, "ML", allele_codes)


def ols_with_covariate(ctx, columns, phen, xxt_eigen_variance_explained=0.75):
    """gwas::ols_with_covariate (src/gwas/ols.rs:278-436) on one GPU: columns f64 [P, n_pools] (the allele columns of
    intercept_and_allele_frequencies[:, 1..]), phen [n_pools, k].  Returns (n_eigenvecs, beta, var, pval) with the
    records as [k, P] arrays."""
    cols = columns
    kin = Kinship(ctx, cols.shape[1], max(1, cols.shape[0]))
    try:
        kin.append_columns(cols)
        kin.gram()
        m = kin.eig_select(cols.shape[0], xxt_eigen_variance_explained)
        beta, var, pval = kin.covar_scan(phen)
        return m, beta, var, pval
    finally:
        kin.close()
