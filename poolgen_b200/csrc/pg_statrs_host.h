// pg_statrs_host.h -- host-side restatement of the statrs 0.16.0 functions behind the reference's p-values
// (StudentsT::cdf -> beta::checked_beta_reg -> gamma::ln_gamma; call sites src/gwas/ols.rs:139,153 and
// src/gwas/correlation_test.rs:65-66), operation for operation, so that the CSV writer can print the very digits the
// reference prints for a given t statistic (the device's per-scan table is smooth to ~1e-12 in ln p, inside the 1e-6
// tolerance of the records but enough to move a 12th printed decimal).  Host libm (glibc) is what Rust's f64::ln /
// exp / sin reach on Linux as well.  Must be compiled WITHOUT floating-point contraction.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace pg {
namespace statrs {

inline double ln_gamma(double x) {
    static const double dk[11] = {2.48574089138753565546e-5,  1.05142378581721974210,     -3.45687097222016235469,
                                  4.51227709466894823700,     -2.98285225323576655721,    1.05639711577126713077,
                                  -1.95428773191645869583e-1, 1.70970543404441224307e-2,  -5.71926117404305781283e-4,
                                  4.63399473359905636708e-6,  -2.71994908488607703910e-9};
    const double R = 10.900511, LN2SEP = 0.6207822376352452223455184457816472122518527279025978;
    const double LNPI = 1.1447298858494001741434273513530587116472948129153, E = 2.71828182845904523536028747135266250;
    const double PI = 3.14159265358979323846264338327950288;
    if (x < 0.5) {
        double s = dk[0];
        for (int i = 1; i < 11; i++) s = s + dk[i] / ((double)i - x);
        return LNPI - log(sin(PI * x)) - log(s) - LN2SEP - (0.5 - x) * log((0.5 - x + R) / E);
    }
    double s = dk[0];
    for (int i = 1; i < 11; i++) s = s + dk[i] / (x + (double)i - 1.0);
    return log(s) + LN2SEP + (x - 0.5) * log((x - 0.5 + R) / E);
}

// approx::ulps_eq!(x, 1.0): epsilon = f64::EPSILON, max_ulps = 4
inline bool ulps_eq_one(double x) {
    if (fabs(x - 1.0) <= DBL_EPSILON) return true;
    if (x < 0.0) return false;
    int64_t a, b;
    const double one = 1.0;
    memcpy(&a, &x, 8);
    memcpy(&b, &one, 8);
    const int64_t d = a > b ? a - b : b - a;
    return d <= 4;
}

// checked_beta_reg(a, b, x) with ln_beta = ln_gamma(a + b) - ln_gamma(a) - ln_gamma(b) formed by the caller in that
// order (it only depends on the degrees of freedom, constant for a whole output file)
inline double beta_reg(double a, double b, double x, double ln_beta) {
    if (!(a > 0.0) || !(b > 0.0) || !(x >= 0.0 && x <= 1.0)) return NAN;
    double bt;
    if (fabs(x) < 1e-10 || ulps_eq_one(x))
        bt = 0.0;
    else
        bt = exp(ln_beta + a * log(x) + b * log(1.0 - x));
    const bool symm = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16, fpmin = DBL_MIN / eps;
    if (symm) {
        const double sw = a;
        x = 1.0 - x;
        a = b;
        b = sw;
    }
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int mi = 1; mi < 141; mi++) {
        const double m = (double)mi, m2 = m * 2.0;
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
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= eps) break;
    }
    return symm ? 1.0 - bt * h / a : bt * h / a;
}

// 2 * (1 - StudentsT(0, 1, df).cdf(|t|)) exactly as src/gwas/ols.rs:153 forms it
struct TwoSidedT {
    double df, ln_beta;
    explicit TwoSidedT(double freedom) : df(freedom) {
        const double a = df / 2.0, b = 0.5;
        ln_beta = ln_gamma(a + b) - ln_gamma(a) - ln_gamma(b);
    }
    inline double operator()(double t_abs) const {
        double cdf;
        if (isinf(df)) {
            cdf = 0.5 * erfc(-t_abs / sqrt(2.0));
        } else {
            const double h = df / (df + t_abs * t_abs);
            const double ib = 0.5 * beta_reg(df / 2.0, 0.5, h, ln_beta);
            cdf = t_abs <= 0.0 ? ib : 1.0 - ib;
        }
        return 2.00 * (1.00 - cdf);
    }
};

}  // namespace statrs
}  // namespace pg
