/*
 * poolgen_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, IEEE f64, no FMA contraction) of the per-locus GWAS
 * scan of jeffersonfparil/poolgen.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (poolgen_b200/, include/poolgen_cuda.h) never links or calls it.
 *
 * Parity status: the reference is a Rust binary crate and no Rust toolchain
 * exists in the build image, so the reference itself could not be run.  This
 * restatement is pinned against every known-answer vector the reference's own
 * unit tests hold for the path (see tests/test_oracle_golden.py):
 *   src/gwas/correlation_test.rs:138-141   r, p and the output line
 *   src/tables/chisq_test.rs:57            output line (chi2=4, p=0.7797774084757156)
 *   src/tables/fisher_exact_test.rs:139-142 log10(5!), hypergeometric ratio, output line
 *   src/base/sync.rs:1557-1632             parse -> counts -> freqs -> filter -> sort
 *   src/gwas/ols.rs:534                    the four betas of the (commented) vector
 *   src/gwas/gwalpha.rs:392-447            the two output lines of gwalpha_ls and of gwalpha_ml: four
 *                                          Nelder-Mead searches over two cost functions, six printed
 *                                          digits each -- they pin the restatement of argmin's solver,
 *                                          statrs' Beta::cdf and bound_parameters_with_logit
 * "parity unpinned" for: MKL inv/det rounding and the ols_iter p-values (no live
 * reference vector exists; df = n-1 per src/gwas/ols.rs:139); mle_iterate (gwas/mle.rs has
 * no known-answer test: its solver is the one test_gwalpha pins, its cost function is
 * checked against the closed-form optimum, tests/test_oracle_golden.py).
 *
 * Third-party arithmetic that is not under /root/reference is restated from the
 * published algorithms: statrs 0.16.0 (ln_gamma, beta_reg, gamma_lr, StudentsT,
 * ChiSquared, Beta), LAPACK dgetrf/dgetri as called by ndarray-linalg 0.16.0, argmin
 * 0.8.1 (NelderMead as prepare_solver_neldermead + Executor::max_iters(1_000) drive it),
 * ndarray 0.15.6 (the summation order of `sum()` / `dot()` on contiguous data).
 *
 * pgo_scan_batch_tight is the "tight" CPU baseline of bench.py: the same arithmetic as
 * the faithful functions (bit-identical records) without the reference's avoidable work.
 */
#ifndef POOLGEN_ORACLE_H
#define POOLGEN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* allele codes follow the sync column order A:T:C:G:N:D (src/base/sync.rs:134-137) */
enum { PGO_A = 0, PGO_T = 1, PGO_C = 2, PGO_G = 3, PGO_N = 4, PGO_D = 5 };
#define PGO_MAX_ALLELES 6

/* FilterStats (src/base/structs_and_traits.rs:69-78); only the fields the sync path reads */
typedef struct {
    int remove_ns;
    uint64_t min_coverage_depth;
    double min_allele_frequency;
    double max_missingness_rate;
    int n_pool_sizes;
    const double *pool_sizes; /* as handed over by the phenotype loader (already normalised) */
} pgo_filter_stats;

/* status codes of the per-locus callbacks */
enum {
    PGO_FILTERED = 0, /* callback returned None because filter() dropped the locus */
    PGO_OK = 1,       /* callback returned Some(line) */
    PGO_FAILED = 2,   /* callback returned None because the regression failed */
    PGO_PANIC = -1    /* the reference would panic (assert / unwrap) */
};

/* ---- statrs 0.16.0 restatement ------------------------------------------------------- */
double pgo_ln_gamma(double x);
double pgo_beta_reg(double a, double b, double x);
double pgo_students_t_cdf(double x, double freedom); /* location 0, scale 1 */
double pgo_gamma_lr(double a, double x);
double pgo_chisq_cdf(double x, double freedom);

/* ---- LAPACK restatement (column-major, n x n) ---------------------------------------- */
int pgo_lu_inverse(double *a, int n);        /* dgetrf + dgetri in place; 0 ok, >0 singular */
double pgo_lu_det(const double *a, int n);   /* dgetrf on a copy, signed product of diag */

/* ---- helpers.rs:103-117 + Rust Display ------------------------------------------------ */
double pgo_sensible_round(double x, int n_digits);
int pgo_f64_to_string(double x, char *buf, size_t cap);
int pgo_round_to_string(double x, int n_digits, char *buf, size_t cap);

/* ---- sync.rs ------------------------------------------------------------------------- */
/* parse one sync line (no trailing newline needed) into n x 6 u64 counts; returns n pools,
 * 0 for a comment line, <0 on malformed input.  chr gets a NUL terminated copy. */
int pgo_parse_sync_line(const char *line, char *chr, size_t chr_cap, uint64_t *pos,
                        uint64_t *counts, int max_pools);
/* LocusCounts::to_frequencies (sync.rs:166-192); counts n x p row-major */
void pgo_to_frequencies(const uint64_t *counts, int n, int p, double *freq);
/* LocusCounts::filter (sync.rs:195-303); in place; returns PGO_OK / PGO_FILTERED / PGO_PANIC */
int pgo_filter(uint64_t *counts, uint8_t *alleles, int n, int *p, const pgo_filter_stats *fs);
/* Sort::sort_by_allele_freq (sync.rs:478-505); in place */
void pgo_sort_by_allele_freq(double *freq, uint8_t *alleles, int n, int p, int decreasing);

/* ---- gwas/ols.rs --------------------------------------------------------------------- */
/* ols() (ols.rs:163-199): X n x p row-major, Y n x k row-major; outputs p x k row-major.
 * returns 0 ok, 1 failed */
int pgo_ols(const double *x, int n, int p, const double *y, int k, double *beta, double *var,
            double *pval, double *tstat);

/* Per-locus numeric result shared by ols_iterate / correlation. Arrays are indexed
 * [allele_out * k + phen]; capacity PGO_MAX_ALLELES * k each (caller allocates). */
typedef struct {
    int status;
    int n_alleles_out;
    uint8_t allele[PGO_MAX_ALLELES];
    double freq_mean[PGO_MAX_ALLELES];
    double *stat; /* beta (ols_iter) or r rounded to 7 digits (pearson_corr) */
    double *var;  /* var(beta) (ols_iter), NaN for pearson_corr */
    double *t;    /* t statistic (ols_iter), NaN for pearson_corr */
    double *pval;
} pgo_locus_result;

/* ols_iterate (ols.rs:201-276); counts n x p row-major (not modified); phen n x k row-major */
int pgo_ols_iterate(const uint64_t *counts, const uint8_t *alleles, int n, int p,
                    const double *phen, int k, const pgo_filter_stats *fs, pgo_locus_result *out);
/* pearsons_correlation (correlation_test.rs:7-71); returns 0 and (r, p) */
int pgo_pearsons_correlation(const double *x, const double *y, int n, double *r, double *pval);
/* correlation (correlation_test.rs:73-129) */
int pgo_correlation(const uint64_t *counts, const uint8_t *alleles, int n, int p,
                    const double *phen, int k, const pgo_filter_stats *fs, pgo_locus_result *out);

/* ---- tables/ ------------------------------------------------------------------------- */
typedef struct {
    int status;
    int n_alleles_out;
    uint8_t allele[PGO_MAX_ALLELES];
    double statistic; /* chi2 (chisq) or p_observed (fisher) */
    double pval;
} pgo_table_result;
int pgo_chisq(const uint64_t *counts, const uint8_t *alleles, int n, int p,
              const pgo_filter_stats *fs, pgo_table_result *out);
double pgo_factorial_log10(double x, int *err);
double pgo_hypergeom_ratio(const double *counts, int n_cells, double log_prod_fac_marginal_sums);
int pgo_fisher(const uint64_t *counts, const uint8_t *alleles, int n, int p,
               const pgo_filter_stats *fs, pgo_table_result *out);

/* ---- argmin 0.8.1 Nelder-Mead as prepare_solver_neldermead + Executor drive it (base/helpers.rs:132-146,
 *      gwas/mle.rs:99-113, gwas/gwalpha.rs:127-137); pinned by the reference's test_gwalpha lines ------------ */
typedef double (*pgo_cost_fn)(const double *x, int d, void *ctx);
int pgo_nelder_mead(pgo_cost_fn cost, void *ctx, int d, double h, int max_iters, double *x_out, double *cost_out);
double pgo_bound_logit(double x, double lower, double upper); /* base/helpers.rs:120-129 */
/* gwalpha_ls (method 0) / gwalpha_ml (method 1), gwas/gwalpha.rs:282-386; phen = gwalpha_fmt matrix n_rows x 3 row-major;
 * out->stat[j] = alpha of allele j */
int pgo_gwalpha(const uint64_t *counts, const uint8_t *alleles, int n, int p, const double *phen, int n_rows,
                int method, const pgo_filter_stats *fs, pgo_locus_result *out);
int pgo_format_gwalpha_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, char *buf, size_t cap);
/* mle_iterate (gwas/mle.rs:232-305): out->stat = beta, var = v_b, t = beta / v_b (sic), pval */
int pgo_mle_iterate(const uint64_t *counts, const uint8_t *alleles, int n, int p, const double *phen, int k,
                    const pgo_filter_stats *fs, pgo_locus_result *out);
/* one regression of mle_with_covariate (gwas/mle.rs:307-463): mle(x, y, false) for one phenotype; x n x p row-major */
int pgo_mle_regress(const double *x, int n, int p, const double *y, double *beta, double *var, double *pval);
int pgo_format_mle_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, int k, char *buf, size_t cap);

/* ---- output lines (ols.rs:255-275, correlation_test.rs:113-128, chisq_test.rs:37-46,
 *      fisher_exact_test.rs:119-129).  Return bytes written (excluding NUL). -------------- */
int pgo_format_ols_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, int k,
                         char *buf, size_t cap);
int pgo_format_corr_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, int k,
                          char *buf, size_t cap);
int pgo_format_chisq_line(const char *chr, uint64_t pos, const pgo_table_result *r, char *buf,
                          size_t cap);
int pgo_format_fisher_line(const char *chr, uint64_t pos, const pgo_table_result *r, char *buf,
                           size_t cap);

/* ---- batch drivers: one OS thread per contiguous locus range (sync.rs:917-939) -------- */
enum { PGO_SCAN_OLS = 0, PGO_SCAN_CORR = 1, PGO_SCAN_CHISQ = 2, PGO_SCAN_FISHER = 3, PGO_SCAN_MLE = 5, PGO_SCAN_GWALPHA_LS = 6,
       PGO_SCAN_GWALPHA_ML = 7 /* gwalpha kinds: phen = gwalpha_fmt matrix, k = its row count, one alpha per allele */ };
/* counts_packed: u32 [locus][allele][pool] (the layout the CUDA library ingests), n_alleles
 * columns named by allele_codes.  Outputs (caller allocated, may be NULL when not wanted):
 *   status[L] (int8), n_out[L] (u8), allele_out[L*6] (u8), freq_mean[L*6],
 *   stat/var/t/pval[L*6*k] indexed [(locus*6 + allele_out)*k + phen]
 * For CHISQ/FISHER k is ignored: stat[L*6*1+0] = statistic, pval[...] = p-value, n_out = #alleles. */
int pgo_scan_batch(int kind, const uint32_t *counts_packed, int64_t n_loci, int n_pools,
                   int n_alleles, const uint8_t *allele_codes, const double *phen, int k,
                   const pgo_filter_stats *fs, int n_threads, int8_t *status, uint8_t *n_out,
                   uint8_t *allele_out, double *freq_mean, double *stat, double *var, double *t,
                   double *pval);

/* the same records, bit for bit, for PGO_SCAN_OLS without the reference's avoidable work: the pool-size total summed
 * once, one inversion of X'X per locus instead of one per phenotype, no per-locus heap allocation ("tight" CPU
 * baseline of BASELINE.md 2; other kinds run the faithful functions) */
int pgo_scan_batch_tight(int kind, const uint32_t *counts_packed, int64_t n_loci, int n_pools,
                         int n_alleles, const uint8_t *allele_codes, const double *phen, int k,
                         const pgo_filter_stats *fs, int n_threads, int8_t *status, uint8_t *n_out,
                         uint8_t *allele_out, double *freq_mean, double *stat, double *var, double *t,
                         double *pval);

#ifdef __cplusplus
}
#endif
/* checker for the ingest kernel's reciprocal-based quotient (see the definition): number of mismatching pairs */
long pgo_check_reciprocal_division(unsigned d_max, long n_random);

#endif
