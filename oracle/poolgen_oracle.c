/*
 * poolgen_oracle.c -- TEST INFRASTRUCTURE ONLY (see poolgen_oracle.h).
 *
 * CPU restatement of poolgen's per-locus GWAS scan.  Every function cites the reference
 * file:line it follows (paths relative to the poolgen repository root).  Compile with
 * -ffp-contract=off: Rust never fuses a*b+c, so neither may this file.
 *
 * Pinning: the reference cannot be built here (no Rust toolchain), so this file is pinned to every known-answer
 * vector the reference's own unit tests hold for the path (tests/test_oracle_golden.py: pearson r and p, chi-square,
 * Fisher, factorial_log10, hypergeom_ratio, sync parse / filter / sort / load, phenotype parse, the betas of the stale
 * ols vector, Rust's f64 Display and rounding).  PARITY UNPINNED where no reference vector exists: the rounding of
 * MKL's inv / det / eig (an LU restatement stands in) and the ols_iter p-value (df = n - 1 read off
 * src/gwas/ols.rs:139).
 *
 * Style note: like the reference, every per-locus step allocates its own temporaries
 * (the reference clones ndarray matrices at each step); this is deliberate so that the
 * timed CPU baseline has the reference's structure, not a hand-optimised one.
 */
#include "poolgen_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define F64_EPSILON 2.220446049250313e-16

/* ====================================================================================== */
/* statrs 0.16.0: function::gamma::ln_gamma (Lanczos, g = 10.900511, 11 coefficients)      */
/* ====================================================================================== */
static const double GAMMA_R = 10.900511;
static const double GAMMA_DK[11] = {
    2.48574089138753565546e-5,  1.05142378581721974210,     -3.45687097222016235469,
    4.51227709466894823700,     -2.98285225323576655721,    1.05639711577126713077,
    -1.95428773191645869583e-1, 1.70970543404441224307e-2,  -5.71926117404305781283e-4,
    4.63399473359905636708e-6,  -2.71994908488607703910e-9,
};
static const double LN_2_SQRT_E_OVER_PI = 0.6207822376352452223455184457816472122518527279025978;
static const double LN_PI = 1.1447298858494001741434273513530587116472948129153;
static const double CONST_E = 2.71828182845904523536028747135266250;
static const double CONST_PI = 3.14159265358979323846264338327950288;

double pgo_ln_gamma(double x) {
    if (x < 0.5) {
        double s = GAMMA_DK[0];
        for (int i = 1; i < 11; i++) s = s + GAMMA_DK[i] / ((double)i - x);
        return LN_PI - log(sin(CONST_PI * x)) - log(s) - LN_2_SQRT_E_OVER_PI -
               (0.5 - x) * log((0.5 - x + GAMMA_R) / CONST_E);
    } else {
        double s = GAMMA_DK[0];
        for (int i = 1; i < 11; i++) s = s + GAMMA_DK[i] / (x + (double)i - 1.0);
        return log(s) + LN_2_SQRT_E_OVER_PI + (x - 0.5) * log((x - 0.5 + GAMMA_R) / CONST_E);
    }
}

/* approx::ulps_eq!(x, 1.0) with the crate defaults (epsilon = f64::EPSILON, max_ulps = 4) */
static int ulps_eq_one(double x) {
    if (fabs(x - 1.0) <= F64_EPSILON) return 1;
    if (x < 0.0) return 0;
    int64_t a, b;
    double one = 1.0;
    memcpy(&a, &x, 8);
    memcpy(&b, &one, 8);
    int64_t d = a > b ? a - b : b - a;
    return d <= 4;
}

/* statrs 0.16.0: function::beta::checked_beta_reg (Lentz continued fraction, <= 140 iters).
 * Call sites in the reference: StudentsT::cdf via gwas/ols.rs:139,153 and
 * gwas/correlation_test.rs:65-66. */
double pgo_beta_reg(double a, double b, double x) {
    if (!(a > 0.0) || !(b > 0.0) || !(x >= 0.0 && x <= 1.0)) return NAN;
    double bt;
    if (fabs(x) < 1e-10 /* prec::is_zero(x, ACC), ACC = 10e-11 */ || ulps_eq_one(x)) {
        bt = 0.0;
    } else {
        bt = exp(pgo_ln_gamma(a + b) - pgo_ln_gamma(a) - pgo_ln_gamma(b) + a * log(x) +
                 b * log(1.0 - x));
    }
    int symm = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16; /* prec::F64_PREC */
    const double fpmin = DBL_MIN / eps;
    if (symm) {
        double swap = a;
        x = 1.0 - x;
        a = b;
        b = swap;
    }
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int mi = 1; mi < 141; mi++) {
        double m = (double)mi;
        double m2 = m * 2.0;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        h = h * d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= eps) break;
    }
    return symm ? 1.0 - bt * h / a : bt * h / a;
}

/* statrs 0.16.0: StudentsT::cdf with location 0, scale 1 */
double pgo_students_t_cdf(double x, double freedom) {
    if (isinf(freedom)) return 0.5 * erfc(-x / sqrt(2.0));
    double k = (x - 0.0) / 1.0;
    double h = freedom / (freedom + k * k);
    double ib = 0.5 * pgo_beta_reg(freedom / 2.0, 0.5, h);
    return x <= 0.0 ? ib : 1.0 - ib;
}

/* statrs 0.16.0: function::gamma::checked_gamma_lr */
double pgo_gamma_lr(double a, double x) {
    if (isnan(a) || isnan(x)) return NAN;
    if (a <= 0.0 || isinf(a)) return NAN;
    if (x <= 0.0 || isinf(x)) return NAN;
    const double eps = 0.000000000000001;
    const double big = 4503599627370496.0;
    const double big_inv = 2.22044604925031308085e-16;
    const double acc = 0.0000000000000011102230246251565; /* prec::DEFAULT_F64_ACC */
    if (fabs(a) < acc) return 1.0;
    if (fabs(x) < acc) return 0.0;
    double ax = a * log(x) - x - pgo_ln_gamma(a);
    if (ax < -709.78271289338399) return a < x ? 1.0 : 0.0;
    if (x <= 1.0 || x <= a) {
        double r2 = a, c2 = 1.0, ans2 = 1.0;
        for (;;) {
            r2 += 1.0;
            c2 *= x / r2;
            ans2 += c2;
            if (c2 / ans2 <= eps) break;
        }
        return exp(ax) * ans2 / a;
    }
    double y = 1.0 - a;
    double z = x + y + 1.0;
    int cnt = 0;
    double p3 = 1.0, q3 = x, p2 = x + 1.0, q2 = z * x;
    double ans = p2 / q2;
    for (;;) {
        y += 1.0;
        z += 2.0;
        cnt += 1;
        double yc = y * (double)cnt;
        double p = p2 * z - p3 * yc;
        double q = q2 * z - q3 * yc;
        p3 = p2;
        p2 = p;
        q3 = q2;
        q2 = q;
        if (fabs(p) > big) {
            p3 *= big_inv;
            p2 *= big_inv;
            q3 *= big_inv;
            q2 *= big_inv;
        }
        if (q != 0.0) {
            double nextans = p / q;
            double error = fabs((ans - nextans) / nextans);
            ans = nextans;
            if (error <= eps) break;
        }
    }
    return 1.0 - exp(ax) * ans;
}

/* statrs 0.16.0: ChiSquared(k) = Gamma(shape k/2, rate 1/2); Gamma::cdf.
 * Call site: tables/chisq_test.rs:33-35 */
double pgo_chisq_cdf(double x, double freedom) {
    double shape = freedom / 2.0, rate = 0.5;
    if (x <= 0.0) return 0.0;
    if (isinf(x)) return 1.0;
    return pgo_gamma_lr(shape, x * rate);
}

/* ====================================================================================== */
/* LAPACK dgetrf (unblocked dgetf2) + dgetri as reached through ndarray-linalg             */
/* `inv()` (gwas/ols.rs:68,77) and `det()` (gwas/ols.rs:72,81).  Column-major n x n.       */
/* ====================================================================================== */
static int lu_factor(double *a, int n, int *ipiv) {
    int info = 0;
    const double sfmin = DBL_MIN;
    for (int j = 0; j < n; j++) {
        int jp = j;
        double amax = fabs(a[j + j * n]);
        for (int i = j + 1; i < n; i++) {
            double v = fabs(a[i + j * n]);
            if (v > amax) {
                amax = v;
                jp = i;
            }
        }
        ipiv[j] = jp;
        if (a[jp + j * n] != 0.0) {
            if (jp != j) {
                for (int l = 0; l < n; l++) {
                    double tmp = a[j + l * n];
                    a[j + l * n] = a[jp + l * n];
                    a[jp + l * n] = tmp;
                }
            }
            if (fabs(a[j + j * n]) >= sfmin) {
                double r = 1.0 / a[j + j * n];
                for (int i = j + 1; i < n; i++) a[i + j * n] = a[i + j * n] * r;
            } else {
                for (int i = j + 1; i < n; i++) a[i + j * n] = a[i + j * n] / a[j + j * n];
            }
        } else if (info == 0) {
            info = j + 1;
        }
        for (int l = j + 1; l < n; l++) {
            double ajl = a[j + l * n];
            for (int i = j + 1; i < n; i++) a[i + l * n] = a[i + l * n] - a[i + j * n] * ajl;
        }
    }
    return info;
}

/* dgetrf + dgetri with caller-provided scratch (ipiv[n], work[n]) */
static int lu_inverse_ws(double *a, int n, int *ipiv, double *work) {
    int info = lu_factor(a, n, ipiv);
    if (info != 0) return info;
    /* dtrti2: inverse of the upper triangle, non-unit diagonal */
    for (int j = 0; j < n; j++) {
        if (a[j + j * n] == 0.0) return j + 1;
    }
    for (int j = 0; j < n; j++) {
        a[j + j * n] = 1.0 / a[j + j * n];
        double ajj = -a[j + j * n];
        /* dtrmv('U','N','N'): x = U[0:j,0:j] * x, x = a[0:j, j] */
        for (int jj = 0; jj < j; jj++) {
            double temp = a[jj + j * n];
            if (temp != 0.0) {
                for (int i = 0; i < jj; i++) a[i + j * n] = a[i + j * n] + temp * a[i + jj * n];
                a[jj + j * n] = a[jj + j * n] * a[jj + jj * n];
            }
        }
        for (int i = 0; i < j; i++) a[i + j * n] = ajj * a[i + j * n];
    }
    /* solve inv(A) * L = inv(U) for inv(A) */
    for (int j = n - 1; j >= 0; j--) {
        for (int i = j + 1; i < n; i++) {
            work[i] = a[i + j * n];
            a[i + j * n] = 0.0;
        }
        if (j < n - 1) {
            for (int l = j + 1; l < n; l++) {
                double temp = -1.0 * work[l];
                if (temp != 0.0)
                    for (int i = 0; i < n; i++) a[i + j * n] = a[i + j * n] + temp * a[i + l * n];
            }
        }
    }
    for (int j = n - 2; j >= 0; j--) {
        int jp = ipiv[j];
        if (jp != j) {
            for (int i = 0; i < n; i++) {
                double tmp = a[i + j * n];
                a[i + j * n] = a[i + jp * n];
                a[i + jp * n] = tmp;
            }
        }
    }
    return 0;
}

int pgo_lu_inverse(double *a, int n) {
    int *ipiv = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    double *work = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int info = lu_inverse_ws(a, n, ipiv, work);
    free(ipiv);
    free(work);
    return info;
}

/* determinant through dgetrf on the scratch copy a[n*n] */
static double lu_det_ws(const double *a_in, int n, double *a, int *ipiv) {
    memcpy(a, a_in, sizeof(double) * (size_t)n * (size_t)n);
    lu_factor(a, n, ipiv); /* singular => a zero on the diagonal => det 0 */
    double det = 1.0;
    for (int j = 0; j < n; j++) {
        det = det * a[j + j * n];
        if (ipiv[j] != j) det = -det;
    }
    return det;
}

double pgo_lu_det(const double *a_in, int n) {
    double *a = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    int *ipiv = (int *)malloc(sizeof(int) * (size_t)n);
    double det = lu_det_ws(a_in, n, a, ipiv);
    free(a);
    free(ipiv);
    return det;
}

/* ====================================================================================== */
/* Rust `f64::to_string()` (Display: shortest round-trip digits, never an exponent) and    */
/* base/helpers.rs:103-117                                                                 */
/* ====================================================================================== */
int pgo_f64_to_string(double x, char *buf, size_t cap) {
    if (isnan(x)) return snprintf(buf, cap, "NaN");
    if (isinf(x)) return snprintf(buf, cap, x > 0 ? "inf" : "-inf");
    if (x == 0.0) return snprintf(buf, cap, signbit(x) ? "-0" : "0");
    char tmp[64];
    int prec;
    for (prec = 0; prec < 17; prec++) {
        snprintf(tmp, sizeof tmp, "%.*e", prec, x);
        if (strtod(tmp, NULL) == x) break;
    }
    /* tmp = [-]d.ddddde[+-]XX */
    char digits[32];
    int nd = 0;
    const char *p = tmp;
    int neg = 0;
    if (*p == '-') {
        neg = 1;
        p++;
    }
    for (; *p && *p != 'e'; p++)
        if (*p != '.') digits[nd++] = *p;
    int exp10 = atoi(p + 1);
    while (nd > 1 && digits[nd - 1] == '0') nd--; /* cannot happen for a minimal precision, but harmless */
    char out[400];
    int o = 0;
    if (neg) out[o++] = '-';
    if (exp10 >= nd - 1) {
        for (int i = 0; i < nd; i++) out[o++] = digits[i];
        for (int i = 0; i < exp10 - (nd - 1); i++) out[o++] = '0';
    } else if (exp10 >= 0) {
        for (int i = 0; i <= exp10; i++) out[o++] = digits[i];
        out[o++] = '.';
        for (int i = exp10 + 1; i < nd; i++) out[o++] = digits[i];
    } else {
        out[o++] = '0';
        out[o++] = '.';
        for (int i = 0; i < -exp10 - 1; i++) out[o++] = '0';
        for (int i = 0; i < nd; i++) out[o++] = digits[i];
    }
    out[o] = 0;
    return snprintf(buf, cap, "%s", out);
}

/* helpers.rs:103-108; factor is the correctly rounded parse of "1e<d>" */
double pgo_sensible_round(double x, int n_digits) {
    char f[16];
    snprintf(f, sizeof f, "1e%d", n_digits);
    double factor = strtod(f, NULL);
    return round(x * factor) / factor; /* C round() = Rust f64::round(): half away from zero */
}

/* helpers.rs:111-117 parse_f64_roundup_and_own */
int pgo_round_to_string(double x, int n_digits, char *buf, size_t cap) {
    int len = pgo_f64_to_string(x, buf, cap);
    if (len < n_digits) return len;
    return pgo_f64_to_string(pgo_sensible_round(x, n_digits), buf, cap);
}

/* ====================================================================================== */
/* base/sync.rs                                                                            */
/* ====================================================================================== */

/* impl Parse<LocusCounts> for String (sync.rs:100-156) */
int pgo_parse_sync_line(const char *line_in, char *chr, size_t chr_cap, uint64_t *pos,
                        uint64_t *counts, int max_pools) {
    size_t len = strlen(line_in);
    char *line = (char *)malloc(len + 1);
    memcpy(line, line_in, len + 1);
    if (len > 0 && line[len - 1] == '\n') {
        line[--len] = 0;
        if (len > 0 && line[len - 1] == '\r') line[--len] = 0;
    }
    if (len == 0 || line[0] == '#') {
        free(line);
        return 0;
    }
    int field = 0, n = 0, rc = 0;
    char *save = NULL;
    /* split on single tabs (str::split keeps empty fields; strtok_r would merge them, so walk by hand) */
    char *cur = line;
    (void)save;
    while (cur) {
        char *tab = strchr(cur, '\t');
        if (tab) *tab = 0;
        if (field == 0) {
            snprintf(chr, chr_cap, "%s", cur);
        } else if (field == 1) {
            char *end = NULL;
            const char *dg = (*cur == '+') ? cur + 1 : cur; /* u64::from_str accepts a leading '+' */
            if (*dg < '0' || *dg > '9') {
                rc = -1;
                break;
            }
            *pos = strtoull(cur, &end, 10);
            if (end == cur || *end != 0) {
                rc = -1;
                break;
            }
        } else if (field >= 3) {
            if (n >= max_pools) {
                rc = -2;
                break;
            }
            /* `.split(":").map(|x| x.parse::<u64>().expect(..))` (sync.rs:144-147): EVERY ':'-separated piece of the
               pool field must parse (optional '+', then digits) or the reference panics; the first six are kept and
               fewer than six is an index panic (sync.rs:148-150) */
            const char *fend = tab ? tab : cur + strlen(cur);
            const char *c = cur;
            int j = 0;
            for (;;) {
                const char *d = (*c == '+') ? c + 1 : c;
                const char *e = d;
                while (e < fend && *e >= '0' && *e <= '9') e++;
                if (e == d || (e < fend && *e != ':')) {
                    rc = -3;
                    break;
                }
                if (j < 6) counts[(size_t)n * 6 + j] = strtoull(d, NULL, 10);
                j++;
                if (e >= fend) break;
                c = e + 1;
            }
            if (!rc && j < 6) rc = -3;
            if (rc) break;
            n++;
        }
        field++;
        cur = tab ? tab + 1 : NULL;
    }
    free(line);
    if (rc) return rc;
    return n;
}

/* LocusCounts::to_frequencies (sync.rs:166-192) */
void pgo_to_frequencies(const uint64_t *counts, int n, int p, double *freq) {
    double *row_sums = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        for (int j = 0; j < p; j++) sum = sum + (double)counts[(size_t)i * p + j];
        row_sums[i] = sum;
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < p; j++)
            freq[(size_t)i * p + j] =
                row_sums[i] == 0.0 ? NAN : (double)counts[(size_t)i * p + j] / row_sums[i];
    free(row_sums);
}

/* ndarray remove_index(Axis(1), j) on a row-major n x p matrix held in a flat buffer */
static void remove_col_u64(uint64_t *m, int n, int p, int j) {
    size_t w = 0;
    for (int i = 0; i < n; i++)
        for (int c = 0; c < p; c++)
            if (c != j) m[w++] = m[(size_t)i * p + c];
}
static void remove_col_f64(double *m, int n, int p, int j) {
    size_t w = 0;
    for (int i = 0; i < n; i++)
        for (int c = 0; c < p; c++)
            if (c != j) m[w++] = m[(size_t)i * p + c];
}

/* LocusCounts::filter (sync.rs:195-303) */
int pgo_filter(uint64_t *counts, uint8_t *alleles, int n, int *p_io, const pgo_filter_stats *fs) {
    int p = *p_io;
    /* Remove Ns (sync.rs:200-213) */
    if (fs->remove_ns) {
        int idx = -1;
        for (int j = 0; j < p; j++)
            if (alleles[j] == PGO_N) {
                idx = j;
                break;
            }
        if (idx != -1) {
            memmove(alleles + idx, alleles + idx + 1, (size_t)(p - idx - 1));
            remove_col_u64(counts, n, p, idx);
            p -= 1;
        }
    }
    /* minimum coverage (sync.rs:217-229) */
    double *sum_coverage = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        for (int j = 0; j < p; j++) sum = sum + (double)counts[(size_t)i * p + j];
        sum_coverage[i] = sum;
    }
    double min_sum_coverage = sum_coverage[0];
    for (int i = 0; i < n; i++)
        if (sum_coverage[i] < min_sum_coverage) min_sum_coverage = sum_coverage[i];
    free(sum_coverage);
    if (min_sum_coverage < (double)fs->min_coverage_depth) {
        *p_io = p;
        return PGO_FILTERED;
    }
    /* minimum allele frequency (sync.rs:238-282) */
    uint64_t *matrix = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)p);
    memcpy(matrix, counts, sizeof(uint64_t) * (size_t)n * (size_t)p);
    double *freq = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
    pgo_to_frequencies(counts, n, p, freq);
    if (n != fs->n_pool_sizes) {
        free(matrix);
        free(freq);
        *p_io = p;
        return PGO_PANIC; /* assert!(n == filter_stats.pool_sizes.len()) sync.rs:254-257 */
    }
    int j = 0;
    while (j < p) {
        double q = 0.0;
        for (int i = 0; i < n; i++) {
            double f = freq[(size_t)i * p + j];
            double term;
            if (isnan(f)) {
                term = 0.0;
            } else {
                double total = 0.0; /* pool_sizes.iter().sum::<f64>() re-summed every time */
                for (int s = 0; s < fs->n_pool_sizes; s++) total = total + fs->pool_sizes[s];
                term = f * (fs->pool_sizes[i] / total);
            }
            q += term;
        }
        if ((q < fs->min_allele_frequency) | (q > (1.00 - fs->min_allele_frequency))) {
            remove_col_f64(freq, n, p, j);
            remove_col_u64(matrix, n, p, j);
            memmove(alleles + j, alleles + j + 1, (size_t)(p - j - 1));
            p -= 1;
        } else {
            j += 1;
        }
    }
    int rc = PGO_OK;
    if (p < 2) {
        rc = PGO_FILTERED;
    } else {
        /* missingness on the first surviving allele column (sync.rs:287-299) */
        int n_missing = 0;
        for (int i = 0; i < n; i++)
            if (isnan(freq[(size_t)i * p + 0])) n_missing += 1;
        if (n_missing == n) rc = PGO_FILTERED;
        else if (((double)n_missing / (double)n) > fs->max_missingness_rate) rc = PGO_FILTERED;
    }
    if (rc == PGO_OK) memcpy(counts, matrix, sizeof(uint64_t) * (size_t)n * (size_t)p); /* sync.rs:301 */
    /* note: on the None paths after the MAF loop the reference has already shortened
     * self.alleles_vector but not self.matrix; callers discard the locus, so only p matters */
    free(matrix);
    free(freq);
    *p_io = p;
    return rc;
}

/* Sort::sort_by_allele_freq (sync.rs:478-505); Rust's sort_by is stable */
void pgo_sort_by_allele_freq(double *freq, uint8_t *alleles, int n, int p, int decreasing) {
    double *column_sums = (double *)malloc(sizeof(double) * (size_t)p);
    int *idx = (int *)malloc(sizeof(int) * (size_t)p);
    for (int j = 0; j < p; j++) {
        double sum = 0.0;
        for (int i = 0; i < n; i++) {
            double v = freq[(size_t)i * p + j];
            if (!isnan(v)) sum = sum + v;
        }
        column_sums[j] = sum;
        idx[j] = j;
    }
    /* stable insertion sort with the reference comparator */
    for (int a = 1; a < p; a++) {
        int v = idx[a];
        int b = a - 1;
        while (b >= 0) {
            int less = decreasing ? (column_sums[v] > column_sums[idx[b]])
                                  : (column_sums[v] < column_sums[idx[b]]);
            if (!less) break;
            idx[b + 1] = idx[b];
            b--;
        }
        idx[b + 1] = v;
    }
    double *sorted = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
    uint8_t sorted_alleles[PGO_MAX_ALLELES];
    for (int c = 0; c < p; c++) {
        for (int i = 0; i < n; i++) sorted[(size_t)i * p + c] = freq[(size_t)i * p + idx[c]];
        sorted_alleles[c] = alleles[idx[c]];
    }
    memcpy(freq, sorted, sizeof(double) * (size_t)n * (size_t)p);
    memcpy(alleles, sorted_alleles, (size_t)p);
    free(sorted);
    free(idx);
    free(column_sums);
}

/* ====================================================================================== */
/* gwas/ols.rs                                                                             */
/* ====================================================================================== */

/* C (m x n) = A (m x k) * B (k x n), all row-major, plain k-ordered accumulation
 * (ndarray `dot`; the reference's matrixmultiply kernel order is reproducible only to rounding) */
static double *matmul(const double *a, int m, int k, const double *b, int n) {
    double *c = (double *)malloc(sizeof(double) * (size_t)m * (size_t)n);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int l = 0; l < k; l++) s = s + a[(size_t)i * k + l] * b[(size_t)l * n + j];
            c[(size_t)i * n + j] = s;
        }
    return c;
}
static double *transpose(const double *a, int m, int n) {
    double *t = (double *)malloc(sizeof(double) * (size_t)m * (size_t)n);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) t[(size_t)j * m + i] = a[(size_t)i * n + j];
    return t;
}

/* UnivariateOrdinaryLeastSquares::{estimate_effects, estimate_variances,
 * estimate_significance} (ols.rs:58-160) for one phenotype column.
 * x is n x p row-major (a private clone, as ols.rs:180 does), y has n entries. */
static int ols_single(const double *x_in, int n, int p, const double *y, double *b_out,
                      double *vb_out, double *t_out, double *pval_out) {
    double *x = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
    memcpy(x, x_in, sizeof(double) * (size_t)n * (size_t)p);
    double *xt = transpose(x, n, p); /* p x n */
    double *b = NULL, *inv_xxt = NULL, *inv_xtx = NULL;
    int fail = 0;
    /* estimate_effects (ols.rs:58-87) */
    if (n < p) {
        inv_xxt = matmul(x, n, p, xt, n); /* n x n */
        if (pgo_lu_inverse(inv_xxt, n) != 0) fail = 1;
        if (!fail && pgo_lu_det(inv_xxt, n) == 0.0) fail = 1;
        if (!fail) {
            double *tmp = matmul(xt, p, n, inv_xxt, n); /* p x n */
            b = matmul(tmp, p, n, y, 1);
            free(tmp);
        }
    } else {
        inv_xtx = matmul(xt, p, n, x, p); /* p x p */
        if (pgo_lu_inverse(inv_xtx, p) != 0) fail = 1;
        if (!fail && pgo_lu_det(inv_xtx, p) == 0.0) fail = 1;
        if (!fail) {
            double *tmp = matmul(inv_xtx, p, p, xt, n); /* p x n */
            b = matmul(tmp, p, n, y, 1);
            free(tmp);
        }
    }
    if (!fail) {
        /* estimate_variances (ols.rs:89-118) */
        double *xb = matmul(x, n, p, b, 1);
        double *e = (double *)malloc(sizeof(double) * (size_t)n);
        for (int i = 0; i < n; i++) e[i] = y[i] - xb[i];
        double ee = 0.0;
        for (int i = 0; i < n; i++) ee = ee + e[i] * e[i];
        double ve = ee / ((double)n - (double)p);
        double *vcv;
        if (n < p) {
            double *t1 = matmul(xt, p, n, inv_xxt, n);
            double *t2 = matmul(t1, p, n, inv_xxt, n);
            vcv = matmul(t2, p, n, x, p);
            for (int i = 0; i < p * p; i++) vcv[i] = ve * vcv[i];
            free(t1);
            free(t2);
        } else {
            vcv = (double *)malloc(sizeof(double) * (size_t)p * (size_t)p);
            for (int i = 0; i < p * p; i++) vcv[i] = ve * inv_xtx[i];
        }
        /* estimate_significance (ols.rs:120-160): StudentsT(0, 1, n - 1) */
        double freedom = (double)n - 1.0;
        if (!(freedom > 0.0)) fail = 2; /* StudentsT::new(...).unwrap() panics */
        for (int i = 0; i < p && !fail; i++) {
            double vb = vcv[(size_t)i * p + i];
            double t = fabs(b[i]) <= F64_EPSILON ? 0.0 : b[i] / sqrt(vb);
            double pv;
            if (fabs(t) <= F64_EPSILON) pv = 1.0;
            else if (isnan(t)) pv = 1.0;
            else pv = 2.00 * (1.00 - pgo_students_t_cdf(fabs(t), freedom));
            b_out[i] = b[i];
            vb_out[i] = vb;
            t_out[i] = t;
            pval_out[i] = pv;
        }
        free(xb);
        free(e);
        free(vcv);
    }
    free(x);
    free(xt);
    free(b);
    free(inv_xxt);
    free(inv_xtx);
    return fail;
}

/* ols() (ols.rs:163-199) */
int pgo_ols(const double *x, int n, int p, const double *y, int k, double *beta, double *var,
            double *pval, double *tstat) {
    double *yj = (double *)malloc(sizeof(double) * (size_t)n);
    double *b = (double *)malloc(sizeof(double) * (size_t)p * 4);
    int fail = 0;
    for (int j = 0; j < k && !fail; j++) {
        for (int i = 0; i < n; i++) yj[i] = y[(size_t)i * k + j];
        fail = ols_single(x, n, p, yj, b, b + p, b + 2 * p, b + 3 * p);
        if (fail) break;
        for (int i = 0; i < p; i++) {
            beta[(size_t)i * k + j] = b[i];
            var[(size_t)i * k + j] = b[p + i];
            if (tstat) tstat[(size_t)i * k + j] = b[2 * p + i];
            pval[(size_t)i * k + j] = b[3 * p + i];
        }
    }
    free(yj);
    free(b);
    return fail;
}

/* RemoveMissing for LocusCountsAndPhenotypes (sync.rs:510-548): pools whose phenotype row
 * mean is NaN are removed from the phenotypes and the counts; returns the new n (0 => Err). */
static int remove_missing(uint64_t *counts, int n, int p, double *phen, int k) {
    int w = 0;
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int j = 0; j < k; j++) s = s + phen[(size_t)i * k + j];
        double mean = s / (double)k;
        if (!isnan(mean)) {
            if (w != i) {
                memcpy(phen + (size_t)w * k, phen + (size_t)i * k, sizeof(double) * (size_t)k);
                memcpy(counts + (size_t)w * p, counts + (size_t)i * p, sizeof(uint64_t) * (size_t)p);
            }
            w++;
        }
    }
    return w;
}

static void result_fill_nan(pgo_locus_result *out, int k) {
    out->n_alleles_out = 0;
    for (int i = 0; i < PGO_MAX_ALLELES; i++) {
        out->allele[i] = 0xff;
        out->freq_mean[i] = NAN;
    }
    for (int i = 0; i < PGO_MAX_ALLELES * k; i++) {
        if (out->stat) out->stat[i] = NAN;
        if (out->var) out->var[i] = NAN;
        if (out->t) out->t[i] = NAN;
        if (out->pval) out->pval[i] = NAN;
    }
}

/* ols_iterate (ols.rs:201-276) */
int pgo_ols_iterate(const uint64_t *counts_in, const uint8_t *alleles_in, int n, int p,
                    const double *phen_in, int k, const pgo_filter_stats *fs,
                    pgo_locus_result *out) {
    result_fill_nan(out, k);
    /* the driver clones counts and the phenotype matrix per locus (sync.rs:858-862) */
    uint64_t *counts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)p);
    double *phen = (double *)malloc(sizeof(double) * (size_t)n * (size_t)k);
    uint8_t alleles[PGO_MAX_ALLELES];
    memcpy(counts, counts_in, sizeof(uint64_t) * (size_t)n * (size_t)p);
    memcpy(phen, phen_in, sizeof(double) * (size_t)n * (size_t)k);
    memcpy(alleles, alleles_in, (size_t)p);
    int status;
    n = remove_missing(counts, n, p, phen, k); /* ols.rs:206 */
    if (n == 0) {
        status = PGO_PANIC;
        goto done;
    }
    status = pgo_filter(counts, alleles, n, &p, fs); /* ols.rs:210-216 */
    if (status != PGO_OK) goto done;
    {
        double *freq = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
        pgo_to_frequencies(counts, n, p, freq);            /* ols.rs:217-220 */
        pgo_sort_by_allele_freq(freq, alleles, n, p, 1);   /* ols.rs:222-225 */
        if (p >= 2) {                                      /* ols.rs:227-230: drops the MAJOR allele */
            remove_col_f64(freq, n, p, 0);
            memmove(alleles, alleles + 1, (size_t)(p - 1));
            p -= 1;
        }
        int px = p + 1; /* intercept, ols.rs:240-246 */
        double *x = (double *)malloc(sizeof(double) * (size_t)n * (size_t)px);
        for (int i = 0; i < n; i++) {
            x[(size_t)i * px] = 1.0;
            for (int j = 1; j < px; j++) x[(size_t)i * px + j] = freq[(size_t)i * p + (j - 1)];
        }
        double *beta = (double *)malloc(sizeof(double) * (size_t)px * (size_t)k * 4);
        double *var = beta + (size_t)px * k, *pv = var + (size_t)px * k, *ts = pv + (size_t)px * k;
        int fail = pgo_ols(x, n, px, phen, k, beta, var, pv, ts); /* ols.rs:249-253 */
        if (fail) {
            status = fail == 2 ? PGO_PANIC : PGO_FAILED;
        } else {
            out->n_alleles_out = p;
            for (int i = 1; i < px; i++) {
                out->allele[i - 1] = alleles[i - 1];
                double s = 0.0; /* x_matrix.column(i).mean() ols.rs:265-268 */
                for (int r = 0; r < n; r++) s = s + x[(size_t)r * px + i];
                out->freq_mean[i - 1] = s / (double)n;
                for (int j = 0; j < k; j++) {
                    out->stat[(size_t)(i - 1) * k + j] = beta[(size_t)i * k + j];
                    out->var[(size_t)(i - 1) * k + j] = var[(size_t)i * k + j];
                    out->t[(size_t)(i - 1) * k + j] = ts[(size_t)i * k + j];
                    out->pval[(size_t)(i - 1) * k + j] = pv[(size_t)i * k + j];
                }
            }
        }
        free(beta);
        free(x);
        free(freq);
    }
done:
    free(counts);
    free(phen);
    out->status = status;
    return status;
}

/* pearsons_correlation (correlation_test.rs:7-71) with method "sensible_corr" */
int pgo_pearsons_correlation(const double *x_in, const double *y_in, int n, double *r_out,
                             double *p_out) {
    double *x = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *y = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; i++)
        if (!isnan(x_in[i]) && !isnan(y_in[i])) {
            x[m] = x_in[i];
            y[m] = y_in[i];
            m++;
        }
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < m; i++) sx = sx + x[i];
    for (int i = 0; i < m; i++) sy = sy + y[i];
    double mu_x = sx / (double)m, mu_y = sy / (double)m;
    double num = 0.0, sxx = 0.0, syy = 0.0;
    for (int i = 0; i < m; i++) {
        double dx = x[i] - mu_x, dy = y[i] - mu_y;
        num = num + dx * dy;
    }
    for (int i = 0; i < m; i++) {
        double dx = x[i] - mu_x;
        sxx = sxx + pow(dx, 2.0);
    }
    for (int i = 0; i < m; i++) {
        double dy = y[i] - mu_y;
        syy = syy + pow(dy, 2.0);
    }
    free(x);
    free(y);
    double denominator = sqrt(sxx) * sqrt(syy);
    double r = num / denominator;
    if (isnan(r)) {
        *r_out = NAN;
        *p_out = NAN;
        return 0;
    }
    double sigma_r_denominator = (1.0 - pow(r, 2.0)) / ((double)n - 2.0);
    if (sigma_r_denominator <= 0.0) {
        *r_out = r; /* unrounded on this path, correlation_test.rs:58-61 */
        *p_out = F64_EPSILON;
        return 0;
    }
    double sigma_r = sqrt(sigma_r_denominator);
    double t = r / sigma_r;
    double pval;
    if (n > 2) pval = 2.00 * (1.00 - pgo_students_t_cdf(fabs(t), (double)n - 2.0));
    else pval = NAN;
    *r_out = pgo_sensible_round(r, 7);
    *p_out = pval;
    return 0;
}

/* correlation (correlation_test.rs:73-129) */
int pgo_correlation(const uint64_t *counts_in, const uint8_t *alleles_in, int n, int p,
                    const double *phen, int k, const pgo_filter_stats *fs, pgo_locus_result *out) {
    result_fill_nan(out, k);
    uint64_t *counts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)p);
    uint8_t alleles[PGO_MAX_ALLELES];
    memcpy(counts, counts_in, sizeof(uint64_t) * (size_t)n * (size_t)p);
    memcpy(alleles, alleles_in, (size_t)p);
    int status = pgo_filter(counts, alleles, n, &p, fs);
    if (status == PGO_OK) {
        double *freq = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
        pgo_to_frequencies(counts, n, p, freq);
        int pk = p >= 2 ? p - 1 : p; /* drop the LAST kept column, correlation_test.rs:94-98 */
        double *xcol = (double *)malloc(sizeof(double) * (size_t)n);
        double *ycol = (double *)malloc(sizeof(double) * (size_t)n);
        out->n_alleles_out = pk;
        for (int i = 0; i < pk; i++) {
            double s = 0.0;
            for (int r = 0; r < n; r++) {
                xcol[r] = freq[(size_t)r * p + i];
                s = s + xcol[r];
            }
            out->allele[i] = alleles[i];
            out->freq_mean[i] = s / (double)n; /* x.mean(), NaN propagates */
            for (int j = 0; j < k; j++) {
                for (int r = 0; r < n; r++) ycol[r] = phen[(size_t)r * k + j];
                double r_, p_;
                pgo_pearsons_correlation(xcol, ycol, n, &r_, &p_);
                out->stat[(size_t)i * k + j] = r_;
                out->pval[(size_t)i * k + j] = p_;
            }
        }
        free(xcol);
        free(ycol);
        free(freq);
    }
    free(counts);
    out->status = status;
    return status;
}

/* ====================================================================================== */
/* tables/chisq_test.rs:5-47                                                               */
/* ====================================================================================== */
int pgo_chisq(const uint64_t *counts_in, const uint8_t *alleles_in, int n, int p,
              const pgo_filter_stats *fs, pgo_table_result *out) {
    uint64_t *counts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)p);
    uint8_t alleles[PGO_MAX_ALLELES];
    memcpy(counts, counts_in, sizeof(uint64_t) * (size_t)n * (size_t)p);
    memcpy(alleles, alleles_in, (size_t)p);
    out->statistic = NAN;
    out->pval = NAN;
    out->n_alleles_out = 0;
    int status = pgo_filter(counts, alleles, n, &p, fs);
    if (status == PGO_OK) {
        double *freq = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
        pgo_to_frequencies(counts, n, p, freq);
        double t = (double)(n * p);
        double total = 0.0;
        for (int i = 0; i < n * p; i++) total = total + freq[i];
        double *row_sums = (double *)calloc((size_t)n, sizeof(double));
        double *col_sums = (double *)calloc((size_t)p, sizeof(double));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < p; j++) row_sums[i] = row_sums[i] + freq[(size_t)i * p + j];
        for (int i = 0; i < n; i++)
            for (int j = 0; j < p; j++) col_sums[j] = col_sums[j] + freq[(size_t)i * p + j];
        double chi2 = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < p; j++) {
                double observed = freq[(size_t)i * p + j];
                double expected = (row_sums[i] * col_sums[j]) / total;
                chi2 += pow(observed - expected, 2.0) / expected;
            }
        out->statistic = chi2;
        out->pval = 1.00 - pgo_chisq_cdf(chi2, t - 1.0);
        out->n_alleles_out = p;
        memcpy(out->allele, alleles, (size_t)p);
        free(row_sums);
        free(col_sums);
        free(freq);
    }
    free(counts);
    out->status = status;
    return status;
}

/* ====================================================================================== */
/* tables/fisher_exact_test.rs                                                             */
/* ====================================================================================== */
/* factorial_log10 (fisher_exact_test.rs:6-18) */
double pgo_factorial_log10(double x, int *err) {
    if (x > 34.0) {
        if (err) *err = 1;
        return NAN;
    }
    double out = 0.0;
    double up = x + 1.0;
    size_t end = up > 0.0 ? (size_t)up : 0; /* `as usize` saturates negatives / NaN to 0 */
    for (size_t i = 2; i < end; i++) out = out + log10((double)i);
    return out;
}

/* hypergeom_ratio (fisher_exact_test.rs:20-30) */
double pgo_hypergeom_ratio(const double *counts, int n_cells, double log_prod_fac_marginal_sums) {
    double prod_fac_sums = 0.0;
    double total = 0.0;
    for (int i = 0; i < n_cells; i++) prod_fac_sums = prod_fac_sums + pgo_factorial_log10(counts[i], NULL);
    for (int i = 0; i < n_cells; i++) total = total + counts[i];
    prod_fac_sums = prod_fac_sums + pgo_factorial_log10(total, NULL);
    return pow(10.0, log_prod_fac_marginal_sums - prod_fac_sums);
}

static double as_usize_f64(double v) { /* `(x) as usize` then back `as f64` */
    if (isnan(v) || v <= 0.0) return 0.0;
    if (v >= 18446744073709551615.0) return 18446744073709551615.0;
    return (double)(uint64_t)v;
}

/* fisher (fisher_exact_test.rs:32-130) */
int pgo_fisher(const uint64_t *counts_in, const uint8_t *alleles_in, int n, int p,
               const pgo_filter_stats *fs, pgo_table_result *out) {
    uint64_t *cu = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)p);
    uint8_t alleles[PGO_MAX_ALLELES];
    memcpy(cu, counts_in, sizeof(uint64_t) * (size_t)n * (size_t)p);
    memcpy(alleles, alleles_in, (size_t)p);
    out->statistic = NAN;
    out->pval = NAN;
    out->n_alleles_out = 0;
    int status = pgo_filter(cu, alleles, n, &p, fs);
    if (status == PGO_OK) {
        int cells = n * p;
        double *c = (double *)malloc(sizeof(double) * (size_t)cells);
        for (int i = 0; i < cells; i++) c[i] = (double)cu[i];
        double total = 0.0;
        for (int i = 0; i < cells; i++) total = total + c[i];
        if (total > 34.0) {
            double coef = 34.0 / total;
            for (int i = 0; i < cells; i++) c[i] = floor(c[i] * coef);
        }
        double *row_sums = (double *)calloc((size_t)n, sizeof(double));
        double *col_sums = (double *)calloc((size_t)p, sizeof(double));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < p; j++) row_sums[i] = row_sums[i] + c[(size_t)i * p + j];
        for (int i = 0; i < n; i++)
            for (int j = 0; j < p; j++) col_sums[j] = col_sums[j] + c[(size_t)i * p + j];
        double lp = 0.0;
        for (int i = 0; i < n; i++) lp = lp + pgo_factorial_log10(row_sums[i], NULL);
        for (int j = 0; j < p; j++) lp = lp + pgo_factorial_log10(col_sums[j], NULL);
        double p_observed = pgo_hypergeom_ratio(c, cells, lp);
        double p_extremes = 0.0;
        int panic = 0;
        for (int max_i = 0; max_i < n && !panic; max_i++) {
            for (int max_j = 0; max_j < p && !panic; max_j++) {
                for (int i = 0; i < n; i++) {
                    for (int j = 0; j < p; j++) {
                        double rs = 0.0, cs = 0.0;
                        for (int jj = 0; jj < j; jj++) rs = rs + c[(size_t)i * p + jj];
                        for (int ii = 0; ii < i; ii++) cs = cs + c[(size_t)ii * p + j];
                        double a = as_usize_f64(row_sums[i] - rs);
                        double b = as_usize_f64(col_sums[j] - cs);
                        double mx = a < b ? a : b;
                        if ((i == (n - 1)) | (j == (p - 1))) c[(size_t)i * p + j] = mx;
                        else if ((i < max_i) | (j < max_j)) c[(size_t)i * p + j] = 0.0;
                        else c[(size_t)i * p + j] = mx;
                    }
                }
                for (int inv_j = 0; inv_j < p; inv_j++) {
                    for (int inv_i = 0; inv_i < n; inv_i++) {
                        int j = p - (inv_j + 1);
                        int i = n - (inv_i + 1);
                        double rs = 0.0, cs = 0.0;
                        for (int jj = 0; jj < p; jj++) rs = rs + c[(size_t)i * p + jj];
                        for (int ii = 0; ii < n; ii++) cs = cs + c[(size_t)ii * p + j];
                        double a = as_usize_f64(row_sums[i] - rs);
                        double b = as_usize_f64(col_sums[j] - cs);
                        double mx = a < b ? a : b;
                        if (mx > 0.0) c[(size_t)i * p + j] = mx;
                    }
                }
                /* assert!(row_sums == counts.sum_axis(Axis(1))) etc. (fisher_exact_test.rs:113-114) */
                for (int i = 0; i < n; i++) {
                    double rs = 0.0;
                    for (int j = 0; j < p; j++) rs = rs + c[(size_t)i * p + j];
                    if (rs != row_sums[i]) panic = 1;
                }
                for (int j = 0; j < p; j++) {
                    double cs = 0.0;
                    for (int i = 0; i < n; i++) cs = cs + c[(size_t)i * p + j];
                    if (cs != col_sums[j]) panic = 1;
                }
                if (panic) break;
                p_extremes += pgo_hypergeom_ratio(c, cells, lp);
            }
        }
        if (panic) {
            status = PGO_PANIC;
        } else {
            out->statistic = p_observed;
            out->pval = p_observed + p_extremes;
            out->n_alleles_out = p;
            memcpy(out->allele, alleles, (size_t)p);
        }
        free(row_sums);
        free(col_sums);
        free(c);
    }
    free(cu);
    out->status = status;
    return status;
}

/* ====================================================================================== */
/* output lines                                                                            */
/* ====================================================================================== */
static const char ALLELE_NAMES[6] = {'A', 'T', 'C', 'G', 'N', 'D'};

int pgo_format_ols_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, int k,
                         char *buf, size_t cap) {
    size_t o = 0;
    char f1[400], f2[400], f3[400];
    buf[0] = 0;
    if (r->status != PGO_OK) return 0;
    for (int i = 0; i < r->n_alleles_out; i++)
        for (int j = 0; j < k; j++) {
            pgo_round_to_string(r->freq_mean[i], 8, f1, sizeof f1);
            pgo_round_to_string(r->stat[(size_t)i * k + j], 6, f2, sizeof f2);
            pgo_round_to_string(r->pval[(size_t)i * k + j], 12, f3, sizeof f3);
            int w = snprintf(buf + o, o < cap ? cap - o : 0, "%s,%llu,%c,%s,Pheno_%d,%s,%s\n", chr,
                             (unsigned long long)pos, ALLELE_NAMES[r->allele[i]], f1, j, f2, f3);
            o += (size_t)w;
        }
    return (int)o;
}

int pgo_format_corr_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, int k,
                          char *buf, size_t cap) {
    size_t o = 0;
    char f1[400], f2[400], f3[400];
    buf[0] = 0;
    if (r->status != PGO_OK) return 0;
    for (int i = 0; i < r->n_alleles_out; i++)
        for (int j = 0; j < k; j++) {
            pgo_f64_to_string(r->freq_mean[i], f1, sizeof f1);
            pgo_round_to_string(r->stat[(size_t)i * k + j], 6, f2, sizeof f2);
            pgo_f64_to_string(r->pval[(size_t)i * k + j], f3, sizeof f3);
            int w = snprintf(buf + o, o < cap ? cap - o : 0, "%s,%llu,%c,%s,Pheno_%d,%s,%s\n", chr,
                             (unsigned long long)pos, ALLELE_NAMES[r->allele[i]], f1, j, f2, f3);
            o += (size_t)w;
        }
    return (int)o;
}

static int join_alleles(const pgo_table_result *r, char *out) {
    int i;
    for (i = 0; i < r->n_alleles_out; i++) out[i] = ALLELE_NAMES[r->allele[i]];
    out[i] = 0;
    return i;
}

int pgo_format_chisq_line(const char *chr, uint64_t pos, const pgo_table_result *r, char *buf,
                          size_t cap) {
    char al[8], f1[400], f2[400];
    buf[0] = 0;
    if (r->status != PGO_OK) return 0;
    join_alleles(r, al);
    pgo_round_to_string(r->statistic, 6, f1, sizeof f1);
    pgo_f64_to_string(r->pval, f2, sizeof f2);
    return snprintf(buf, cap, "%s,%llu,%s,%s,%s\n", chr, (unsigned long long)pos, al, f1, f2);
}

int pgo_format_fisher_line(const char *chr, uint64_t pos, const pgo_table_result *r, char *buf,
                           size_t cap) {
    char al[8], f1[400], f2[400];
    buf[0] = 0;
    if (r->status != PGO_OK) return 0;
    join_alleles(r, al);
    pgo_f64_to_string(r->statistic, f1, sizeof f1);
    pgo_f64_to_string(r->pval, f2, sizeof f2);
    return snprintf(buf, cap, "%s,%llu,%s,%s,%s\n", chr, (unsigned long long)pos, al, f1, f2);
}

/* ====================================================================================== */
/* argmin 0.8.1 Nelder-Mead (solver/neldermead/mod.rs) as the reference drives it:          */
/* `prepare_solver_neldermead(p, h)` (base/helpers.rs:132-146) builds p + 1 vertices of     */
/* dimension p -- vertex i is h everywhere and h + 0.5 at coordinate i, the last vertex all  */
/* h -- and `Executor::new(cost, solver).configure(|s| s.max_iters(1_000)).run()`            */
/* (gwas/mle.rs:99-113, gwas/gwalpha.rs:127-137).  The crate is not under /root/reference;   */
/* this restates its published algorithm: alpha = 1, gamma = 2, rho = sigma = 0.5,           */
/* sd_tolerance = f64::EPSILON; every iteration reflects the worst vertex through the        */
/* centroid of the others; f_best <= f_r < f_second_worst accepts the reflection, f_r <      */
/* f_best tries the expansion and keeps the better of the two, f_r >= f_second_worst         */
/* contracts -- outside (towards the reflected point, accepted when f_c <= f_r) when         */
/* f_r < f_worst, inside (towards the worst vertex, accepted when f_c < f_worst) otherwise    */
/* -- and a failed contraction shrinks every vertex half way towards the best one;           */
/* vertices are kept sorted by cost (stable);                                               */
/* the run stops after 1,000 iterations or when the sample standard deviation of the         */
/* vertex costs drops below the tolerance; the result is the best vertex.                   */
/* PINNED by the reference's own test_gwalpha lines (gwas/gwalpha.rs:392-447), which go      */
/* through this solver with both cost functions: tests/test_oracle_golden.py.               */
/* ====================================================================================== */
#define PGO_NM_MAXD 16
typedef double (*pgo_cost_fn)(const double *x, int d, void *ctx);

static void nm_sort(double v[][PGO_NM_MAXD], double *c, int nv, int d) {
    /* stable insertion sort by cost; partial_cmp -> Equal for NaN */
    for (int a = 1; a < nv; a++) {
        double ca = c[a], va[PGO_NM_MAXD];
        memcpy(va, v[a], sizeof(double) * (size_t)d);
        int b = a - 1;
        while (b >= 0 && ca < c[b]) {
            c[b + 1] = c[b];
            memcpy(v[b + 1], v[b], sizeof(double) * (size_t)d);
            b--;
        }
        c[b + 1] = ca;
        memcpy(v[b + 1], va, sizeof(double) * (size_t)d);
    }
}

/* returns the number of iterations run; x_out = best vertex */
int pgo_nelder_mead(pgo_cost_fn cost, void *ctx, int d, double h, int max_iters, double *x_out, double *cost_out) {
    double v[PGO_NM_MAXD + 1][PGO_NM_MAXD], c[PGO_NM_MAXD + 1];
    const int nv = d + 1;
    for (int i = 0; i < nv; i++)
        for (int j = 0; j < d; j++) v[i][j] = (i == j) ? h + 0.5 : h;
    for (int i = 0; i < nv; i++) c[i] = cost(v[i], d, ctx);
    nm_sort(v, c, nv, d);
    int it = 0;
    for (;;) {
        /* terminate(): sample standard deviation of the costs */
        double c0 = 0.0;
        for (int i = 0; i < nv; i++) c0 = c0 + c[i];
        c0 = c0 / (double)nv;
        double ss = 0.0;
        for (int i = 0; i < nv; i++) ss = ss + (c[i] - c0) * (c[i] - c0);
        const double sd = sqrt(1.0 / ((double)nv - 1.0) * ss);
        if (sd < F64_EPSILON) break;
        if (it >= max_iters) break;
        double x0[PGO_NM_MAXD], xr[PGO_NM_MAXD], xt[PGO_NM_MAXD];
        /* centroid of all vertices but the worst */
        for (int j = 0; j < d; j++) x0[j] = v[0][j];
        for (int i = 1; i < nv - 1; i++)
            for (int j = 0; j < d; j++) x0[j] = x0[j] + v[i][j];
        const double inv = 1.0 / (double)(nv - 1);
        for (int j = 0; j < d; j++) x0[j] = x0[j] * inv;
        /* reflection: x0 + alpha (x0 - x_worst) */
        for (int j = 0; j < d; j++) xr[j] = x0[j] + (x0[j] - v[nv - 1][j]) * 1.0;
        const double fr = cost(xr, d, ctx);
        if (fr < c[nv - 2] && fr >= c[0]) {
            memcpy(v[nv - 1], xr, sizeof(double) * (size_t)d);
            c[nv - 1] = fr;
        } else if (fr < c[0]) {
            for (int j = 0; j < d; j++) xt[j] = x0[j] + (xr[j] - x0[j]) * 2.0; /* expansion */
            const double fe = cost(xt, d, ctx);
            if (fe < fr) {
                memcpy(v[nv - 1], xt, sizeof(double) * (size_t)d);
                c[nv - 1] = fe;
            } else {
                memcpy(v[nv - 1], xr, sizeof(double) * (size_t)d);
                c[nv - 1] = fr;
            }
        } else if (fr >= c[nv - 2]) {
            int shrink = 0;
            if (fr < c[nv - 1]) {
                /* outside contraction: between the centroid and the reflected point */
                for (int j = 0; j < d; j++) xt[j] = x0[j] + (xr[j] - x0[j]) * 0.5;
                const double fc = cost(xt, d, ctx);
                if (fc <= fr) {
                    memcpy(v[nv - 1], xt, sizeof(double) * (size_t)d);
                    c[nv - 1] = fc;
                } else {
                    shrink = 1;
                }
            } else {
                /* inside contraction: between the centroid and the worst vertex */
                for (int j = 0; j < d; j++) xt[j] = x0[j] + (v[nv - 1][j] - x0[j]) * 0.5;
                const double fc = cost(xt, d, ctx);
                if (fc < c[nv - 1]) {
                    memcpy(v[nv - 1], xt, sizeof(double) * (size_t)d);
                    c[nv - 1] = fc;
                } else {
                    shrink = 1;
                }
            }
            if (shrink) { /* every vertex moves half way towards the best one */
                for (int i = 1; i < nv; i++) {
                    for (int j = 0; j < d; j++) v[i][j] = v[0][j] + (v[i][j] - v[0][j]) * 0.5;
                    c[i] = cost(v[i], d, ctx);
                }
            }
        } else {
            /* only reachable with NaN costs */
            for (int i = 1; i < nv; i++) {
                for (int j = 0; j < d; j++) v[i][j] = v[0][j] + (v[i][j] - v[0][j]) * 0.5;
                c[i] = cost(v[i], d, ctx);
            }
        }
        nm_sort(v, c, nv, d);
        it++;
    }
    memcpy(x_out, v[0], sizeof(double) * (size_t)d);
    if (cost_out) *cost_out = c[0];
    return it;
}

/* bound_parameters_with_logit (base/helpers.rs:120-129) */
double pgo_bound_logit(double x, double lower, double upper) { return lower + ((upper - lower) / (1.00 + exp(-x))); }

/* ndarray 0.15 numeric_util::unrolled_fold with + (what `.sum()` does on contiguous data): eight partial sums */
static double ndarray_sum_contig(const double *xs, int len) {
    double acc = 0.0, p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    while (len >= 8) {
        p0 = p0 + xs[0]; p1 = p1 + xs[1]; p2 = p2 + xs[2]; p3 = p3 + xs[3];
        p4 = p4 + xs[4]; p5 = p5 + xs[5]; p6 = p6 + xs[6]; p7 = p7 + xs[7];
        xs += 8;
        len -= 8;
    }
    acc = acc + (p0 + p4);
    acc = acc + (p1 + p5);
    acc = acc + (p2 + p6);
    acc = acc + (p3 + p7);
    for (int i = 0; i < len && i < 7; i++) acc = acc + xs[i];
    return acc;
}
/* numeric_util::unrolled_dot */
static double ndarray_dot_contig(const double *xs, const double *ys, int len) {
    double sum = 0.0, p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    while (len >= 8) {
        p0 = p0 + xs[0] * ys[0]; p1 = p1 + xs[1] * ys[1]; p2 = p2 + xs[2] * ys[2]; p3 = p3 + xs[3] * ys[3];
        p4 = p4 + xs[4] * ys[4]; p5 = p5 + xs[5] * ys[5]; p6 = p6 + xs[6] * ys[6]; p7 = p7 + xs[7] * ys[7];
        xs += 8;
        ys += 8;
        len -= 8;
    }
    sum = sum + (p0 + p4);
    sum = sum + (p1 + p5);
    sum = sum + (p2 + p6);
    sum = sum + (p3 + p7);
    for (int i = 0; i < len && i < 7; i++) sum = sum + xs[i] * ys[i];
    return sum;
}

/* statrs 0.16.0 Beta::cdf */
static double beta_cdf(double a, double b, double x) {
    if (x < 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    if (isinf(a)) return x < 1.0 ? 0.0 : 1.0;
    if (isinf(b)) return 1.0;
    if (ulps_eq_one(a) && ulps_eq_one(b)) return x;
    return pgo_beta_reg(a, b, x);
}

/* ---- gwas/gwalpha.rs ---------------------------------------------------------------------- */
#define GWALPHA_LO F64_EPSILON
#define GWALPHA_HI 10.00
typedef struct {
    int n;
    const double *percs_a, *percs_b, *percs_a0, *percs_b0, *q_prime;
} gwalpha_ctx;

/* least_squares_beta (gwalpha.rs:11-41) */
static double gwalpha_cost_ls(const double *x, int d, void *vctx) {
    const gwalpha_ctx *g = (const gwalpha_ctx *)vctx;
    double sh[4];
    for (int i = 0; i < 4; i++) sh[i] = pgo_bound_logit(x[i], GWALPHA_LO, GWALPHA_HI);
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < g->n; i++) {
        const double da = g->percs_a[i] - beta_cdf(sh[0], sh[1], g->q_prime[i]);
        const double db = g->percs_b[i] - beta_cdf(sh[2], sh[3], g->q_prime[i]);
        sa += pow(da, 2.0);
        sb += pow(db, 2.0);
    }
    return sa + sb;
}
/* maximum_likelihood_beta (gwalpha.rs:43-82) */
static double gwalpha_cost_ml(const double *x, int d, void *vctx) {
    const gwalpha_ctx *g = (const gwalpha_ctx *)vctx;
    double sh[4];
    for (int i = 0; i < 4; i++) sh[i] = pgo_bound_logit(x[i], GWALPHA_LO, GWALPHA_HI);
    double la = 0.0, lb = 0.0;
    for (int i = 0; i < g->n; i++) {
        double da = beta_cdf(sh[0], sh[1], g->percs_a[i]) - beta_cdf(sh[0], sh[1], g->percs_a0[i]);
        double db = beta_cdf(sh[2], sh[3], g->percs_b[i]) - beta_cdf(sh[2], sh[3], g->percs_b0[i]);
        if (da < F64_EPSILON) da = F64_EPSILON;
        if (db < F64_EPSILON) db = F64_EPSILON;
        la += log10(da);
        lb += log10(db);
    }
    return -la - lb;
}

/* gwalpha_ls / gwalpha_ml (gwalpha.rs:282-386).  phen: the gwalpha_fmt matrix, n_rows x 3 row-major (column 0 bins,
 * column 1 q, column 2 = sig, min, max, then -inf).  Outputs: out->allele / freq_mean (rounded to 6 digits by the line
 * formatter) and out->stat[j] = alpha of allele j (k = 1).  method 0 = LS, 1 = ML. */
int pgo_gwalpha(const uint64_t *counts_in, const uint8_t *alleles_in, int n, int p, const double *phen,
                int n_rows, int method, const pgo_filter_stats *fs, pgo_locus_result *out) {
    result_fill_nan(out, 1);
    uint64_t *counts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)p);
    uint8_t alleles[PGO_MAX_ALLELES];
    memcpy(counts, counts_in, sizeof(uint64_t) * (size_t)n * (size_t)p);
    memcpy(alleles, alleles_in, (size_t)p);
    int status = pgo_filter(counts, alleles, n, &p, fs);
    if (status != PGO_OK) {
        free(counts);
        return out->status = status;
    }
    double *freq = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
    pgo_to_frequencies(counts, n, p, freq);
    pgo_sort_by_allele_freq(freq, alleles, n, p, 1);
    if (p >= 2) {
        remove_col_f64(freq, n, p, 0);
        memmove(alleles, alleles + 1, (size_t)(p - 1));
        p -= 1;
    }
    /* bins and q without the -inf fillers; sig, min, max (gwalpha.rs:197-216) */
    double *bins = (double *)malloc(sizeof(double) * (size_t)n_rows * 8);
    double *q = bins + n_rows, *qp = q + n_rows, *ba = qp + n_rows, *bb = ba + n_rows, *pa = bb + n_rows,
           *pb = pa + n_rows, *fa = pb + n_rows;
    int m = 0, mq = 0;
    for (int i = 0; i < n_rows; i++) {
        if (phen[(size_t)i * 3 + 0] != -INFINITY) bins[m++] = phen[(size_t)i * 3 + 0];
        if (phen[(size_t)i * 3 + 1] != -INFINITY) q[mq++] = phen[(size_t)i * 3 + 1];
    }
    const double sig = phen[0 * 3 + 2], min = phen[1 * 3 + 2], max = phen[2 * 3 + 2];
    if (n != m || n_rows < 3) {
        free(bins);
        free(freq);
        free(counts);
        return out->status = PGO_FILTERED; /* None */
    }
    double *pa0 = (double *)malloc(sizeof(double) * (size_t)n * 2), *pb0 = pa0 + n;
    out->n_alleles_out = p;
    for (int j = 0; j < p; j++) {
        /* prepare_freqs_and_qprime (gwalpha.rs:225-280) */
        for (int i = 0; i < n; i++) fa[i] = freq[(size_t)i * p + j];
        double p_a;
        if (p == 1) {
            p_a = ndarray_dot_contig(fa, bins, n); /* a one-column matrix: the column view is contiguous */
        } else {
            p_a = 0.0;
            for (int i = 0; i < n; i++) p_a += fa[i] * bins[i];
        }
        qp[0] = 0.0;
        for (int i = 1; i < n; i++) qp[i] = (q[i] - min) / (max - min);
        for (int i = 0; i < n; i++) {
            ba[i] = (fa[i]) * bins[i] / (p_a);
            bb[i] = (1.0 - fa[i]) * bins[i] / (1.0 - p_a);
        }
        pa[0] = ba[0];
        pb[0] = bb[0];
        for (int i = 1; i < n; i++) {
            pa[i] = ndarray_sum_contig(ba, i + 1);
            pb[i] = ndarray_sum_contig(bb, i + 1);
        }
        pa0[0] = pb0[0] = 0.0;
        for (int i = 0; i < n - 1; i++) {
            pa0[i + 1] = pa[i];
            pb0[i + 1] = pb[i];
        }
        gwalpha_ctx g = {n, pa, pb, pa0, pb0, qp};
        double sol[4];
        pgo_nelder_mead(method == 0 ? gwalpha_cost_ls : gwalpha_cost_ml, &g, 4, 1.0, 1000, sol, NULL);
        for (int i = 0; i < 4; i++) sol[i] = pgo_bound_logit(sol[i], GWALPHA_LO, GWALPHA_HI);
        const double a_mu = min + (max - min) * (sol[0] / (sol[0] + sol[1]));
        const double b_mu = min + (max - min) * (sol[2] / (sol[2] + sol[3]));
        const double alpha = (2.00 * sqrt(p_a * (1.0 - p_a))) * (a_mu - b_mu) / sig;
        out->allele[j] = alleles[j];
        out->freq_mean[j] = (p == 1 ? ndarray_sum_contig(fa, n) : ({ double s_ = 0.0; for (int i = 0; i < n; i++) s_ = s_ + fa[i]; s_; })) / (double)n;
        out->stat[j] = alpha;
    }
    free(pa0);
    free(bins);
    free(freq);
    free(counts);
    return out->status = PGO_OK;
}

/* ---- gwas/mle.rs -------------------------------------------------------------------------- */
typedef struct {
    int n, p;
    const double *x; /* n x p row-major */
    const double *y; /* n */
} mle_ctx;

/* negative_likelihood_normal_distribution_sigma_and_beta (mle.rs:13-30); params = [logit sigma2, betas] */
static double mle_cost(const double *par, int d, void *vctx) {
    const mle_ctx *m = (const mle_ctx *)vctx;
    const double sigma2 = pgo_bound_logit(par[0], F64_EPSILON, 1e9);
    double ss = 0.0;
    for (int i = 0; i < m->n; i++) {
        /* x.dot(&betas): one contiguous row.dot per element (sequential below 8 columns, ndarray's 8 partial sums from there) */
        const double xb = ndarray_dot_contig(m->x + (size_t)i * m->p, par + 1, m->p);
        const double e = m->y[i] - xb;
        ss = ss + pow(e, 2.0);
    }
    return ((double)m->n / 2.00) * log(2.00 * 3.14159265358979323846264338327950288 * sigma2) + (1.00 / sigma2) * ss;
}

/* remove_collinearities_in_x (mle.rs:56-83), literally: returns the new column count, or -1 where the reference's
 * `i -= 1` underflows (it would index out of bounds and panic) */
static int mle_remove_collinear(double *x, int n, int p) {
    if (p == 2) return p;
    long i = 1;
    while (i < p) {
        long j = i + 1;
        while (j < p) {
            if (i < 0) return -1;
            double *ci = (double *)malloc(sizeof(double) * (size_t)n * 2), *cj = ci + n;
            for (int r = 0; r < n; r++) {
                ci[r] = x[(size_t)r * p + i];
                cj[r] = x[(size_t)r * p + j];
            }
            double cor = 0.0, pv = NAN;
            if (pgo_pearsons_correlation(ci, cj, n, &cor, &pv) != 0) cor = 0.0;
            free(ci);
            if (fabs(cor) >= 0.99) {
                remove_col_f64(x, n, p, (int)j);
                p -= 1;
                i -= 1;
                j -= 1;
            }
            j += 1;
        }
        i += 1;
    }
    return p;
}

/* mle_iterate (mle.rs:232-305) with mle() (193-230) and the Regression impl (85-191).  out->stat = beta,
 * out->var = v_b = ve diag((X'X)^-1), out->t = beta / v_b (sic, mle.rs:176), out->pval; rows of removed collinear
 * columns keep the reference's zeros */
int pgo_mle_iterate(const uint64_t *counts_in, const uint8_t *alleles_in, int n, int p, const double *phen, int k,
                    const pgo_filter_stats *fs, pgo_locus_result *out) {
    result_fill_nan(out, k);
    uint64_t *counts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)p);
    uint8_t alleles[PGO_MAX_ALLELES];
    memcpy(counts, counts_in, sizeof(uint64_t) * (size_t)n * (size_t)p);
    memcpy(alleles, alleles_in, (size_t)p);
    int status = pgo_filter(counts, alleles, n, &p, fs);
    if (status != PGO_OK) {
        free(counts);
        return out->status = status;
    }
    double *freq = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
    pgo_to_frequencies(counts, n, p, freq);
    pgo_sort_by_allele_freq(freq, alleles, n, p, 1);
    if (p >= 2) {
        remove_col_f64(freq, n, p, 0);
        memmove(alleles, alleles + 1, (size_t)(p - 1));
        p -= 1;
    }
    const int px = p + 1;
    double *x = (double *)malloc(sizeof(double) * (size_t)n * (size_t)px * 2), *xw = x + (size_t)n * px;
    for (int i = 0; i < n; i++) {
        x[(size_t)i * px] = 1.0;
        for (int j = 1; j < px; j++) x[(size_t)i * px + j] = freq[(size_t)i * p + (j - 1)];
    }
    double *yj = (double *)malloc(sizeof(double) * (size_t)n);
    double beta[PGO_MAX_ALLELES + 1][64], vb[PGO_MAX_ALLELES + 1][64], tt[PGO_MAX_ALLELES + 1][64], pv[PGO_MAX_ALLELES + 1][64];
    int fail = k > 64;
    for (int i = 0; i < px && !fail; i++)
        for (int j = 0; j < k; j++) beta[i][j] = vb[i][j] = pv[i][j] = 0.0, tt[i][j] = NAN; /* Array2::zeros */
    for (int j = 0; j < k && !fail; j++) {
        memcpy(xw, x, sizeof(double) * (size_t)n * (size_t)px);
        const int pw = mle_remove_collinear(xw, n, px);
        if (pw < 0) {
            fail = 2;
            break;
        }
        for (int i = 0; i < n; i++) yj[i] = phen[(size_t)i * k + j];
        mle_ctx m = {n, pw, xw, yj};
        double par[PGO_NM_MAXD];
        pgo_nelder_mead(mle_cost, &m, pw + 1, 1.0, 1000, par, NULL);
        const double ve = pgo_bound_logit(par[0], F64_EPSILON, 1e9);
        /* estimate_variances (mle.rs:116-155) */
        double *xt = transpose(xw, n, pw), *vcv = NULL;
        if (n < pw) {
            double *inv = matmul(xw, n, pw, xt, n);
            if (pgo_lu_inverse(inv, n) != 0 || pgo_lu_det(inv, n) == 0.0) fail = 1;
            if (!fail) {
                double *t1 = matmul(xt, pw, n, inv, n), *t2 = matmul(t1, pw, n, inv, n);
                vcv = matmul(t2, pw, n, xw, pw);
                for (int i = 0; i < pw * pw; i++) vcv[i] = ve * vcv[i];
                free(t1);
                free(t2);
            }
            free(inv);
        } else {
            double *inv = matmul(xt, pw, n, xw, pw);
            if (pgo_lu_inverse(inv, pw) != 0 || pgo_lu_det(inv, pw) == 0.0) fail = 1;
            if (!fail) {
                vcv = (double *)malloc(sizeof(double) * (size_t)pw * (size_t)pw);
                for (int i = 0; i < pw * pw; i++) vcv[i] = ve * inv[i];
            }
            free(inv);
        }
        free(xt);
        if (fail) break;
        const double freedom = (double)n - 1.0;
        if (!(freedom > 0.0)) {
            fail = 2;
            free(vcv);
            break;
        }
        for (int i = 0; i < pw; i++) {
            const double b = par[1 + i], v = vcv[(size_t)i * pw + i];
            const double t = b / v; /* mle.rs:176: the variance, not its square root */
            double pval;
            if (isinf(t)) pval = 0.0;
            else if (isnan(t)) pval = 1.0;
            else pval = 2.00 * (1.00 - pgo_students_t_cdf(fabs(t), freedom));
            beta[i][j] = b;
            vb[i][j] = v;
            tt[i][j] = t;
            pv[i][j] = pval;
        }
        free(vcv);
    }
    if (fail) {
        status = fail == 2 ? PGO_PANIC : PGO_FAILED;
    } else {
        out->n_alleles_out = p;
        for (int i = 1; i < px; i++) {
            out->allele[i - 1] = alleles[i - 1];
            double sm = 0.0;
            for (int r = 0; r < n; r++) sm = sm + x[(size_t)r * px + i];
            out->freq_mean[i - 1] = sm / (double)n;
            for (int j = 0; j < k; j++) {
                out->stat[(size_t)(i - 1) * k + j] = beta[i][j];
                out->var[(size_t)(i - 1) * k + j] = vb[i][j];
                out->t[(size_t)(i - 1) * k + j] = tt[i][j];
                out->pval[(size_t)(i - 1) * k + j] = pv[i][j];
            }
        }
    }
    free(yj);
    free(x);
    free(freq);
    free(counts);
    return out->status = status;
}

/* mle() with remove_collinearities = false and one phenotype, as mle_with_covariate calls it for X = [1 | PCs | g]
 * (mle.rs:193-230 through the Regression impl 85-191; call site 370-392): beta / var / pval of all p columns; nonzero
 * where the reference returns Err (the caller then records NaN) */
int pgo_mle_regress(const double *x, int n, int p, const double *y, double *beta, double *var, double *pval) {
    if (p + 1 > PGO_NM_MAXD) return -2;
    mle_ctx m = {n, p, x, y};
    double par[PGO_NM_MAXD];
    pgo_nelder_mead(mle_cost, &m, p + 1, 1.0, 1000, par, NULL);
    const double ve = pgo_bound_logit(par[0], F64_EPSILON, 1e9);
    double *xt = transpose(x, n, p), *vcv = NULL;
    int fail = 0;
    if (n < p) {
        double *inv = matmul(x, n, p, xt, n);
        if (pgo_lu_inverse(inv, n) != 0 || pgo_lu_det(inv, n) == 0.0) fail = 1;
        if (!fail) {
            double *t1 = matmul(xt, p, n, inv, n), *t2 = matmul(t1, p, n, inv, n);
            vcv = matmul(t2, p, n, x, p);
            for (int i = 0; i < p * p; i++) vcv[i] = ve * vcv[i];
            free(t1);
            free(t2);
        }
        free(inv);
    } else {
        double *inv = matmul(xt, p, n, x, p);
        if (pgo_lu_inverse(inv, p) != 0 || pgo_lu_det(inv, p) == 0.0) fail = 1;
        if (!fail) {
            vcv = (double *)malloc(sizeof(double) * (size_t)p * (size_t)p);
            for (int i = 0; i < p * p; i++) vcv[i] = ve * inv[i];
        }
        free(inv);
    }
    free(xt);
    if (fail) return 1;
    const double freedom = (double)n - 1.0;
    if (!(freedom > 0.0)) { /* StudentsT::new(..).unwrap() panics */
        free(vcv);
        return 2;
    }
    for (int i = 0; i < p; i++) {
        const double b = par[1 + i], v = vcv[(size_t)i * p + i];
        const double t = b / v; /* mle.rs:176 */
        double pv;
        if (isinf(t)) pv = 0.0;
        else if (isnan(t)) pv = 1.0;
        else pv = 2.00 * (1.00 - pgo_students_t_cdf(fabs(t), freedom));
        beta[i] = b;
        var[i] = v;
        pval[i] = pv;
    }
    free(vcv);
    return 0;
}

/* output lines of mle_iterate (mle.rs:283-303): beta rounded to 6 digits, p-value unrounded */
int pgo_format_mle_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, int k, char *buf, size_t cap) {
    size_t w = 0;
    buf[0] = 0;
    if (r->status != PGO_OK) return 0;
    for (int i = 0; i < r->n_alleles_out; i++)
        for (int j = 0; j < k; j++) {
            char f[420], b[420], pvs[420];
            pgo_round_to_string(r->freq_mean[i], 8, f, sizeof f);
            pgo_round_to_string(r->stat[(size_t)i * k + j], 6, b, sizeof b);
            pgo_f64_to_string(r->pval[(size_t)i * k + j], pvs, sizeof pvs);
            int len = snprintf(buf + w, cap - w, "%s,%llu,%c,%s,Pheno_%d,%s,%s\n", chr, (unsigned long long)pos,
                               "ATCGND"[r->allele[i]], f, j, b, pvs);
            if (len < 0 || (size_t)len >= cap - w) return -1;
            w += (size_t)len;
        }
    return (int)w;
}

/* output lines of gwalpha_ls / gwalpha_ml (gwalpha.rs:320-331): chr,pos,allele,freq r6,Pheno_0,alpha r6,Unknown */
int pgo_format_gwalpha_lines(const char *chr, uint64_t pos, const pgo_locus_result *r, char *buf, size_t cap) {
    size_t w = 0;
    buf[0] = 0;
    if (r->status != PGO_OK) return 0;
    for (int j = 0; j < r->n_alleles_out; j++) {
        char f[420], a[420];
        pgo_round_to_string(r->freq_mean[j], 6, f, sizeof f);
        pgo_round_to_string(r->stat[j], 6, a, sizeof a);
        int len = snprintf(buf + w, cap - w, "%s,%llu,%c,%s,Pheno_0,%s,Unknown\n", chr, (unsigned long long)pos,
                           "ATCGND"[r->allele[j]], f, a);
        if (len < 0 || (size_t)len >= cap - w) return -1;
        w += (size_t)len;
    }
    return (int)w;
}

/* ====================================================================================== */
/* "tight" ols_iterate: the SAME arithmetic, operation for operation, as pgo_ols_iterate   */
/* (results are bit-identical, tests/test_oracle_golden.py), without the reference's       */
/* avoidable work -- the honest upper bound of what its CPU path could do: */
/*   - the pool-size total of sync.rs:262-270 is summed once per batch, not per pool and   */
/*     allele of every locus (the value is the same every time);                          */
/*   - X'X is inverted (and det(inv) checked) once per locus, not once per phenotype       */
/*     (ols.rs:180 clones X and refactors it for each trait); (X'X)^-1 X' is formed once;  */
/*   - no per-locus heap allocation: one workspace per thread.                            */
/* Loci whose phenotype rows contain NaN take the faithful function.                       */
/* ====================================================================================== */
typedef struct {
    uint64_t *counts; /* n x 6 */
    double *f1;       /* n x 6 first-stage frequencies */
    double *freq;     /* n x 6 renormalised, then sorted */
    double *x;        /* n x 6 */
    double *tmp;      /* 6 x n  (X'X)^-1 X' */
    double *e;        /* n */
    double *wnorm;    /* n: s_i / total */
} tight_ws;

static size_t tight_ws_doubles(int n) { return (size_t)n * (6 + 6 + 6 + 6 + 1 + 1); }

static int ols_iterate_tight(const uint32_t *packed /* [A][n] */, const uint8_t *alleles_in, int n, int A,
                             const double *phen, int k, const pgo_filter_stats *fs, tight_ws *ws,
                             pgo_locus_result *out) {
    result_fill_nan(out, k);
    out->status = PGO_FILTERED;
    /* columns in play, in order (remove_index keeps the order of the others) */
    int col[PGO_MAX_ALLELES], p = 0;
    uint8_t alle[PGO_MAX_ALLELES];
    for (int j = 0; j < A; j++) {
        if (fs->remove_ns && alleles_in[j] == PGO_N) continue; /* sync.rs:200-213 */
        col[p] = j;
        alle[p] = alleles_in[j];
        p++;
    }
    /* coverage (sync.rs:217-229) and first-stage frequencies (sync.rs:242, 166-192) in one pass */
    double min_cov = 0.0;
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        for (int j = 0; j < p; j++) sum = sum + (double)packed[(size_t)col[j] * n + i];
        if (i == 0 || sum < min_cov) min_cov = sum;
        for (int j = 0; j < p; j++)
            ws->f1[(size_t)i * 6 + j] = sum == 0.0 ? NAN : (double)packed[(size_t)col[j] * n + i] / sum;
    }
    if (min_cov < (double)fs->min_coverage_depth) return PGO_FILTERED;
    if (n != fs->n_pool_sizes) return out->status = PGO_PANIC;
    /* MAF (sync.rs:258-282) */
    int kept[PGO_MAX_ALLELES], pk = 0;
    for (int j = 0; j < p; j++) {
        double q = 0.0;
        for (int i = 0; i < n; i++) {
            double f = ws->f1[(size_t)i * 6 + j];
            double term = isnan(f) ? 0.0 : f * ws->wnorm[i];
            q += term;
        }
        if (!((q < fs->min_allele_frequency) | (q > (1.00 - fs->min_allele_frequency)))) kept[pk++] = j;
    }
    if (pk < 2) return PGO_FILTERED;
    {
        int n_missing = 0; /* sync.rs:287-299 */
        for (int i = 0; i < n; i++)
            if (isnan(ws->f1[(size_t)i * 6 + kept[0]])) n_missing += 1;
        if (n_missing == n) return PGO_FILTERED;
        if (((double)n_missing / (double)n) > fs->max_missingness_rate) return PGO_FILTERED;
    }
    /* to_frequencies over the kept alleles (ols.rs:217-220) and the column sums of the sort (sync.rs:478-505) */
    double csum[PGO_MAX_ALLELES];
    for (int a = 0; a < pk; a++) csum[a] = 0.0;
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        for (int a = 0; a < pk; a++) sum = sum + (double)packed[(size_t)col[kept[a]] * n + i];
        for (int a = 0; a < pk; a++) {
            double v = sum == 0.0 ? NAN : (double)packed[(size_t)col[kept[a]] * n + i] / sum;
            ws->freq[(size_t)i * 6 + a] = v;
            if (!isnan(v)) csum[a] = csum[a] + v;
        }
    }
    int idx[PGO_MAX_ALLELES];
    for (int a = 0; a < pk; a++) idx[a] = a;
    for (int a = 1; a < pk; a++) { /* stable, decreasing */
        int v = idx[a], b = a - 1;
        while (b >= 0 && csum[v] > csum[idx[b]]) {
            idx[b + 1] = idx[b];
            b--;
        }
        idx[b + 1] = v;
    }
    /* drop the major allele (ols.rs:227-230), X = [1 | freqs] (ols.rs:240-246) */
    const int pa = pk - 1, px = pk;
    for (int i = 0; i < n; i++) {
        ws->x[(size_t)i * px] = 1.0;
        for (int j = 1; j < px; j++) ws->x[(size_t)i * px + j] = ws->freq[(size_t)i * 6 + idx[j]];
    }
    if (n < px) return -2; /* the n < p branch: the caller takes the faithful function */
    /* X'X, its inverse and det(inv) once per locus (ols.rs:76-84) */
    double xtx[36], det_scratch[36], work[6];
    int ipiv[6];
    for (int i = 0; i < px; i++)
        for (int j = 0; j < px; j++) {
            double s = 0.0;
            for (int l = 0; l < n; l++) s = s + ws->x[(size_t)l * px + i] * ws->x[(size_t)l * px + j];
            xtx[i * px + j] = s;
        }
    if (lu_inverse_ws(xtx, px, ipiv, work) != 0) return out->status = PGO_FAILED;
    if (lu_det_ws(xtx, px, det_scratch, ipiv) == 0.0) return out->status = PGO_FAILED;
    for (int i = 0; i < px; i++)
        for (int r = 0; r < n; r++) {
            double s = 0.0;
            for (int l = 0; l < px; l++) s = s + xtx[i * px + l] * ws->x[(size_t)r * px + l];
            ws->tmp[(size_t)i * n + r] = s;
        }
    const double freedom = (double)n - 1.0;
    if (!(freedom > 0.0)) return out->status = PGO_PANIC;
    for (int j = 0; j < k; j++) {
        double b[6];
        for (int i = 0; i < px; i++) {
            double s = 0.0;
            for (int r = 0; r < n; r++) s = s + ws->tmp[(size_t)i * n + r] * phen[(size_t)r * k + j];
            b[i] = s;
        }
        double ee = 0.0;
        for (int r = 0; r < n; r++) {
            double xb = 0.0;
            for (int l = 0; l < px; l++) xb = xb + ws->x[(size_t)r * px + l] * b[l];
            ws->e[r] = phen[(size_t)r * k + j] - xb;
        }
        for (int r = 0; r < n; r++) ee = ee + ws->e[r] * ws->e[r];
        const double ve = ee / ((double)n - (double)px);
        for (int i = 1; i < px; i++) {
            const double vb = ve * xtx[i * px + i];
            const double t = fabs(b[i]) <= F64_EPSILON ? 0.0 : b[i] / sqrt(vb);
            double pv;
            if (fabs(t) <= F64_EPSILON) pv = 1.0;
            else if (isnan(t)) pv = 1.0;
            else pv = 2.00 * (1.00 - pgo_students_t_cdf(fabs(t), freedom));
            out->stat[(size_t)(i - 1) * k + j] = b[i];
            out->var[(size_t)(i - 1) * k + j] = vb;
            out->t[(size_t)(i - 1) * k + j] = t;
            out->pval[(size_t)(i - 1) * k + j] = pv;
        }
    }
    out->n_alleles_out = pa;
    for (int i = 1; i < px; i++) {
        out->allele[i - 1] = alle[kept[idx[i]]];
        double s = 0.0;
        for (int r = 0; r < n; r++) s = s + ws->x[(size_t)r * px + i];
        out->freq_mean[i - 1] = s / (double)n;
    }
    return out->status = PGO_OK;
}

/* ====================================================================================== */
/* batch driver: contiguous locus ranges, one OS thread each (sync.rs:917-939)             */
/* ====================================================================================== */
typedef struct {
    int kind;
    const uint32_t *counts_packed;
    int64_t lo, hi;
    int n_pools, n_alleles, k;
    const uint8_t *allele_codes;
    const double *phen;
    const pgo_filter_stats *fs;
    int8_t *status;
    uint8_t *n_out;
    uint8_t *allele_out;
    double *freq_mean, *stat, *var, *t, *pval;
    int tight; /* ols_iter only: the allocation-free variant (ols_iterate_tight) */
    int phen_rows; /* gwalpha: rows of the gwalpha_fmt matrix */
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *jb = (batch_job *)arg;
    int n = jb->n_pools, A = jb->n_alleles, k = jb->k > 0 ? jb->k : 1;
    if (jb->phen_rows) k = 1; /* gwalpha: one alpha per allele; jb->k carried the rows of the gwalpha_fmt matrix */
    size_t cap = (size_t)PGO_MAX_ALLELES * (size_t)k;
    double *scratch = (double *)malloc(sizeof(double) * cap * 4);
    if (jb->tight && jb->kind == PGO_SCAN_OLS) {
        int phen_nan = 0;
        for (size_t i = 0; i < (size_t)n * (size_t)k; i++)
            if (isnan(jb->phen[i])) phen_nan = 1;
        double *buf = (double *)malloc(sizeof(double) * tight_ws_doubles(n));
        tight_ws ws;
        ws.f1 = buf;
        ws.freq = buf + (size_t)n * 6;
        ws.x = buf + (size_t)n * 12;
        ws.tmp = buf + (size_t)n * 18;
        ws.e = buf + (size_t)n * 24;
        ws.wnorm = buf + (size_t)n * 25;
        ws.counts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)A);
        double total = 0.0; /* sync.rs:262-270, once */
        for (int s_ = 0; s_ < jb->fs->n_pool_sizes; s_++) total = total + jb->fs->pool_sizes[s_];
        for (int i = 0; i < n && i < jb->fs->n_pool_sizes; i++) ws.wnorm[i] = jb->fs->pool_sizes[i] / total;
        for (int64_t l = jb->lo; l < jb->hi; l++) {
            const uint32_t *src = jb->counts_packed + (size_t)l * (size_t)A * (size_t)n;
            pgo_locus_result r;
            r.stat = scratch;
            r.var = scratch + cap;
            r.t = scratch + 2 * cap;
            r.pval = scratch + 3 * cap;
            int status = phen_nan ? -2 : ols_iterate_tight(src, jb->allele_codes, n, A, jb->phen, k, jb->fs, &ws, &r);
            if (status == -2) { /* NaN phenotypes or n < p: the faithful function */
                for (int a = 0; a < A; a++)
                    for (int i = 0; i < n; i++) ws.counts[(size_t)i * A + a] = src[(size_t)a * n + i];
                status = pgo_ols_iterate(ws.counts, jb->allele_codes, n, A, jb->phen, k, jb->fs, &r);
            }
            if (jb->freq_mean)
                memcpy(jb->freq_mean + (size_t)l * PGO_MAX_ALLELES, r.freq_mean, sizeof(double) * PGO_MAX_ALLELES);
            if (jb->stat) memcpy(jb->stat + (size_t)l * cap, r.stat, sizeof(double) * cap);
            if (jb->var) memcpy(jb->var + (size_t)l * cap, r.var, sizeof(double) * cap);
            if (jb->t) memcpy(jb->t + (size_t)l * cap, r.t, sizeof(double) * cap);
            if (jb->pval) memcpy(jb->pval + (size_t)l * cap, r.pval, sizeof(double) * cap);
            if (jb->status) jb->status[l] = (int8_t)status;
            if (jb->n_out) jb->n_out[l] = (uint8_t)(status == PGO_OK ? r.n_alleles_out : 0);
            if (jb->allele_out) memcpy(jb->allele_out + (size_t)l * PGO_MAX_ALLELES, r.allele, PGO_MAX_ALLELES);
        }
        free(ws.counts);
        free(buf);
        free(scratch);
        return NULL;
    }
    for (int64_t l = jb->lo; l < jb->hi; l++) {
        /* "parse": build the n x p u64 LocusCounts matrix the callback receives */
        uint64_t *counts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * (size_t)A);
        const uint32_t *src = jb->counts_packed + (size_t)l * (size_t)A * (size_t)n;
        for (int a = 0; a < A; a++)
            for (int i = 0; i < n; i++) counts[(size_t)i * A + a] = src[(size_t)a * n + i];
        int status, nout = 0;
        uint8_t alle[PGO_MAX_ALLELES];
        memset(alle, 0xff, sizeof alle);
        if (jb->kind == PGO_SCAN_OLS || jb->kind == PGO_SCAN_CORR || jb->kind >= PGO_SCAN_MLE) {
            pgo_locus_result r;
            r.stat = scratch;
            r.var = scratch + cap;
            r.t = scratch + 2 * cap;
            r.pval = scratch + 3 * cap;
            if (jb->kind == PGO_SCAN_OLS)
                status = pgo_ols_iterate(counts, jb->allele_codes, n, A, jb->phen, k, jb->fs, &r);
            else if (jb->kind == PGO_SCAN_MLE)
                status = pgo_mle_iterate(counts, jb->allele_codes, n, A, jb->phen, k, jb->fs, &r);
            else if (jb->kind == PGO_SCAN_GWALPHA_LS || jb->kind == PGO_SCAN_GWALPHA_ML)
                status = pgo_gwalpha(counts, jb->allele_codes, n, A, jb->phen, jb->phen_rows,
                                     jb->kind == PGO_SCAN_GWALPHA_ML, jb->fs, &r);
            else
                status = pgo_correlation(counts, jb->allele_codes, n, A, jb->phen, k, jb->fs, &r);
            nout = r.n_alleles_out;
            memcpy(alle, r.allele, PGO_MAX_ALLELES);
            if (jb->freq_mean)
                memcpy(jb->freq_mean + (size_t)l * PGO_MAX_ALLELES, r.freq_mean,
                       sizeof(double) * PGO_MAX_ALLELES);
            if (jb->stat) memcpy(jb->stat + (size_t)l * cap, r.stat, sizeof(double) * cap);
            if (jb->var) memcpy(jb->var + (size_t)l * cap, r.var, sizeof(double) * cap);
            if (jb->t) memcpy(jb->t + (size_t)l * cap, r.t, sizeof(double) * cap);
            if (jb->pval) memcpy(jb->pval + (size_t)l * cap, r.pval, sizeof(double) * cap);
        } else {
            pgo_table_result r;
            if (jb->kind == PGO_SCAN_CHISQ) status = pgo_chisq(counts, jb->allele_codes, n, A, jb->fs, &r);
            else status = pgo_fisher(counts, jb->allele_codes, n, A, jb->fs, &r);
            nout = r.n_alleles_out;
            if (status == PGO_OK) memcpy(alle, r.allele, (size_t)nout);
            if (jb->stat) jb->stat[(size_t)l * cap] = r.statistic;
            if (jb->pval) jb->pval[(size_t)l * cap] = r.pval;
        }
        if (jb->status) jb->status[l] = (int8_t)status;
        if (jb->n_out) jb->n_out[l] = (uint8_t)(status == PGO_OK ? nout : 0);
        if (jb->allele_out) memcpy(jb->allele_out + (size_t)l * PGO_MAX_ALLELES, alle, PGO_MAX_ALLELES);
        free(counts);
    }
    free(scratch);
    return NULL;
}

static int scan_batch_impl(int kind, int tight, const uint32_t *counts_packed, int64_t n_loci, int n_pools,
                           int n_alleles, const uint8_t *allele_codes, const double *phen, int k,
                           const pgo_filter_stats *fs, int n_threads, int8_t *status, uint8_t *n_out,
                           uint8_t *allele_out, double *freq_mean, double *stat, double *var, double *t,
                           double *pval) {
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_loci) n_threads = n_loci > 0 ? (int)n_loci : 1;
    if (n_alleles > PGO_MAX_ALLELES || n_alleles < 1) return -1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * (size_t)n_threads);
    for (int i = 0; i < n_threads; i++) {
        batch_job jb = {kind, counts_packed, n_loci * i / n_threads, n_loci * (i + 1) / n_threads,
                        n_pools, n_alleles, k, allele_codes, phen, fs, status, n_out, allele_out,
                        freq_mean, stat, var, t, pval, tight, (kind == PGO_SCAN_GWALPHA_LS || kind == PGO_SCAN_GWALPHA_ML) ? k : 0};
        jobs[i] = jb;
        if (n_threads == 1) batch_worker(&jobs[i]);
        else pthread_create(&th[i], NULL, batch_worker, &jobs[i]);
    }
    if (n_threads > 1)
        for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    free(jobs);
    return 0;
}

int pgo_scan_batch(int kind, const uint32_t *counts_packed, int64_t n_loci, int n_pools,
                   int n_alleles, const uint8_t *allele_codes, const double *phen, int k,
                   const pgo_filter_stats *fs, int n_threads, int8_t *status, uint8_t *n_out,
                   uint8_t *allele_out, double *freq_mean, double *stat, double *var, double *t,
                   double *pval) {
    return scan_batch_impl(kind, 0, counts_packed, n_loci, n_pools, n_alleles, allele_codes, phen, k, fs,
                           n_threads, status, n_out, allele_out, freq_mean, stat, var, t, pval);
}

int pgo_scan_batch_tight(int kind, const uint32_t *counts_packed, int64_t n_loci, int n_pools,
                         int n_alleles, const uint8_t *allele_codes, const double *phen, int k,
                         const pgo_filter_stats *fs, int n_threads, int8_t *status, uint8_t *n_out,
                         uint8_t *allele_out, double *freq_mean, double *stat, double *var, double *t,
                         double *pval) {
    return scan_batch_impl(kind, 1, counts_packed, n_loci, n_pools, n_alleles, allele_codes, phen, k, fs,
                           n_threads, status, n_out, allele_out, freq_mean, stat, var, t, pval);
}


/* ====================================================================================== */
/* Checker for one device-side shortcut (not a restatement of the reference): the ingest   */
/* kernel forms c / d as RN(q + (c - q d) r) with r = RN(1 / d), q = RN(c r) -- Markstein's */
/* correction step -- instead of one IEEE division per allele.  Every operation is an      */
/* IEEE-754 round-to-nearest operation, so the host's fma() reproduces the device bit for   */
/* bit: returns how many (c, d) pairs differ from the plain quotient -- all c <= d <= d_max, */
/* then n_random pairs of 32-bit operands.                                                 */
/* ====================================================================================== */
long pgo_check_reciprocal_division(unsigned d_max, long n_random) {
    long bad = 0;
    for (uint32_t d = 1; d <= d_max; d++) {
        const double dd = (double)d, r = 1.0 / dd;
        for (uint32_t c = 0; c <= d; c++) {
            const double q = (double)c * r;
            if (fma(fma(-q, dd, (double)c), r, q) != (double)c / dd) bad++;
        }
    }
    uint64_t s = 88172645463325252ull;
    for (long i = 0; i < n_random; i++) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        uint32_t d = (uint32_t)(s >> 32), c = (uint32_t)s;
        if (!d) d = 1;
        if (i & 1) c %= d;
        const double dd = (double)d, r = 1.0 / dd, q = (double)c * r;
        if (fma(fma(-q, dd, (double)c), r, q) != (double)c / dd) bad++;
    }
    return bad;
}
