// pg_tables.cu -- the count-table callbacks: tables::chisq (src/tables/chisq_test.rs:5-47) and
// tables::fisher (src/tables/fisher_exact_test.rs:32-130), one thread per locus.  The raw u32 counts
// [locus][allele][pool] of a CTA's loci are staged through shared memory with coalesced 128-bit loads;
// each thread then walks its own locus in the reference's exact sequential order, so the keep-mask
// needs no rounding analysis here.
#include <stdlib.h>

#include "pg_device.cuh"
#include "pg_internal.h"

namespace pg {

constexpr int kTabThreads = 128;
constexpr int kTabMaxPools = 16;
constexpr int kTabMaxCells = kTabMaxPools * PG_MAX_ALLELES;

__constant__ double c_lf10[36];  // log10(x!) accumulated like factorial_log10 (fisher_exact_test.rs:6-18)

// LocusCounts::filter on one locus held as cnt[j*n + i]; returns status and the kept column list
__device__ int table_filter(const uint32_t *cnt, const TableParams &p, int *cols, int &pk) {
    const int n = p.n;
    int col[PG_MAX_ALLELES];
    int pc = 0;
    for (int j = 0; j < p.A_in; j++)
        if (j != p.drop_col) col[pc++] = j;
    // depth per pool = sequential f64 sum over the remaining columns (exact: integers)
    double dmin = 0.0;
    for (int i = 0; i < n; i++) {
        double d = 0.0;
        for (int a = 0; a < pc; a++) d += (double)cnt[col[a] * n + i];
        if (i == 0 || d < dmin) dmin = d;
    }
    pk = 0;
    if (dmin < p.min_depth_f) return PG_LOCUS_FILTERED;
    for (int a = 0; a < pc; a++) {
        double q = 0.0;
        for (int i = 0; i < n; i++) {
            double d = 0.0;
            for (int b = 0; b < pc; b++) d += (double)cnt[col[b] * n + i];
            const double f = (d == 0.0) ? nan("") : (double)cnt[col[a] * n + i] / d;
            const double term = (f != f) ? 0.0 : __dmul_rn(f, p.w[i]);
            q = __dadd_rn(q, term);
        }
        if (!((q < p.maf) | (q > p.one_minus_maf))) cols[pk++] = col[a];
    }
    if (pk < 2) return PG_LOCUS_FILTERED;
    int miss = 0;
    for (int i = 0; i < n; i++) {
        double d = 0.0;
        for (int b = 0; b < pc; b++) d += (double)cnt[col[b] * n + i];
        if (d == 0.0) miss++;
    }
    if (miss == n) return PG_LOCUS_FILTERED;
    if (((double)miss / (double)n) > p.max_miss) return PG_LOCUS_FILTERED;
    return PG_LOCUS_OK;
}

__device__ __forceinline__ double as_usize_f64(double v) {  // `(x) as usize as f64`
    if (v != v || v <= 0.0) return 0.0;
    if (v >= 18446744073709551615.0) return 18446744073709551615.0;
    return (double)(unsigned long long)v;
}

__device__ double hypergeom_ratio_dev(const double *c, int cells, double lp) {
    double s = 0.0, total = 0.0;
    for (int i = 0; i < cells; i++) s = s + c_lf10[(int)c[i]];
    for (int i = 0; i < cells; i++) total = total + c[i];
    s = s + c_lf10[(int)total];
    return pow(10.0, lp - s);
}


// ---- register-resident variant for the common small tables (NP pools x NA stored alleles, C5: 2 x 4..6) ------------
// Same arithmetic in the same order as the generic kernel below, but the table lives in registers (every index is a
// compile-time constant, removed alleles stay in place as zero cells under a kept mask, which preserves the order of
// every sum over the kept columns), the chi-square tail uses the finite closed forms of the regularised gamma function
// for integer and half-integer shape, and 10^x is exp10.

// upper tail Q(a, x) of the regularised gamma function for 2a = df integer:
//   a = m      : Q = e^-x sum_{j<m} x^j / j!
//   a = m + 1/2: Q = erfc(sqrt x) + e^-x sum_{j<m} x^(j+1/2) / Gamma(j + 3/2)
// (finite sums of positive terms: accurate to a few ulp; statrs reaches the same value through its series / continued
// fraction to 1e-15).  The caller forms p = 1 - (1 - Q) like the reference's 1 - cdf.
__device__ __forceinline__ double gamma_q_halfint(int df, double x) {
    const double ex = exp(-x);
    if (df & 1) {
        const int m = df >> 1;
        const double sx = sqrt(x);
        double term = 1.1283791670955125738961589031215 * sx;  // 2 sqrt(x) / sqrt(pi) = x^(1/2) / Gamma(3/2)
        double sum = 0.0;
        for (int j = 0; j < m; j++) {
            sum += term;
            term *= x / ((double)j + 1.5);
        }
        return erfc(sx) + ex * sum;
    }
    const int m = df >> 1;
    double term = 1.0, sum = 0.0;
    for (int j = 0; j < m; j++) {
        sum += term;
        term *= x / (double)(j + 1);
    }
    return ex * sum;
}

// lf: the log10-factorial table in SHARED memory -- the lookups are indexed by the cell values, which differ from lane
// to lane, and a divergent index serialises constant-cache reads
template <int NP, int NA>
__device__ __forceinline__ double ratio_t(const double (&c)[NP][NA], unsigned km, double lp, const double *lf) {
    double s = 0.0, total = 0.0;
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int j = 0; j < NA; j++)
            if ((km >> j) & 1u) s = s + lf[(int)c[i][j]];
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int j = 0; j < NA; j++)
            if ((km >> j) & 1u) total = total + c[i][j];
    s = s + lf[(int)total];
    return exp10(lp - s);
}

template <int NP, int NA>
__device__ __forceinline__ double ratio_i(const int (&c)[NP][NA], unsigned km, double lp, const double *lf) {
    double s = 0.0;
    int total = 0;
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int j = 0; j < NA; j++)
            if ((km >> j) & 1u) {
                s = s + lf[c[i][j]];
                total += c[i][j];
            }
    s = s + lf[min(total, 35)];
    return exp10(lp - s);
}

template <int NP, int NA>
__global__ void __launch_bounds__(kTabThreads, (NP * NA <= 8) ? 12 : 8) tables_kernel_t(const TableParams p) {
    extern __shared__ __align__(16) uint32_t sm_counts[];
    __shared__ double s_lf[36];
    if (threadIdx.x < 36) s_lf[threadIdx.x] = c_lf10[threadIdx.x];
    constexpr int per_locus = NA * NP;
    const int64_t tiles = (p.n_loci + kTabThreads - 1) / kTabThreads;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t l0 = tile * kTabThreads;
        const int nl = (int)min((int64_t)kTabThreads, p.n_loci - l0);
        const size_t words = (size_t)nl * per_locus;
        const uint32_t *src = p.counts + (size_t)l0 * per_locus;
        __syncthreads();
        const size_t w4 = words / 4;
        for (size_t i = threadIdx.x; i < w4; i += kTabThreads)
            reinterpret_cast<uint4 *>(sm_counts)[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
        for (size_t i = w4 * 4 + threadIdx.x; i < words; i += kTabThreads) sm_counts[i] = src[i];
        __syncthreads();
        if ((int)threadIdx.x >= nl) continue;
        const int64_t locus = l0 + threadIdx.x;
        double cnt[NP][NA];  // cnt[i][j] = count of pool i, stored allele j (0 for the dropped N column)
        unsigned present = 0;  // stored columns that take part (N dropped when remove_ns)
#pragma unroll
        for (int j = 0; j < NA; j++) {
            const bool on = j != p.drop_col;
            if (on) present |= 1u << j;
#pragma unroll
            for (int i = 0; i < NP; i++) cnt[i][j] = on ? (double)sm_counts[(size_t)threadIdx.x * per_locus + j * NP + i] : 0.0;
        }
        // LocusCounts::filter (src/base/sync.rs:195-303), sequential sums in the reference's order
        double dep[NP];
        double dmin = 0.0;
        int miss = 0;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            double d = 0.0;
#pragma unroll
            for (int j = 0; j < NA; j++)
                if ((present >> j) & 1u) d += cnt[i][j];
            dep[i] = d;
            if (i == 0 || d < dmin) dmin = d;
            if (d == 0.0) miss++;
        }
        int status = PG_LOCUS_OK;
        unsigned km = 0;
        if (dmin < p.min_depth_f) {
            status = PG_LOCUS_FILTERED;
        } else {
#pragma unroll
            for (int j = 0; j < NA; j++) {
                if (!((present >> j) & 1u)) continue;
                double q = 0.0;
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    const double f = (dep[i] == 0.0) ? nan("") : cnt[i][j] / dep[i];
                    const double term = (f != f) ? 0.0 : __dmul_rn(f, p.w[i]);
                    q = __dadd_rn(q, term);
                }
                if (!((q < p.maf) | (q > p.one_minus_maf))) km |= 1u << j;
            }
            if (__popc(km) < 2)
                status = PG_LOCUS_FILTERED;
            else if (miss == NP || ((double)miss / (double)NP) > p.max_miss)
                status = PG_LOCUS_FILTERED;
        }
        const int pk = __popc(km);
        double stat = nan(""), pval = nan("");
        if (status == PG_LOCUS_OK) {
            double c[NP][NA], rs[NP], cs[NA];
            if (p.kind == PG_KIND_CHISQ) {
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    double d = 0.0;
#pragma unroll
                    for (int j = 0; j < NA; j++)
                        if ((km >> j) & 1u) d = d + cnt[i][j];
#pragma unroll
                    for (int j = 0; j < NA; j++) c[i][j] = ((km >> j) & 1u) ? ((d == 0.0) ? nan("") : cnt[i][j] / d) : 0.0;
                }
                double total = 0.0;
#pragma unroll
                for (int i = 0; i < NP; i++)
#pragma unroll
                    for (int j = 0; j < NA; j++)
                        if ((km >> j) & 1u) total = total + c[i][j];
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    rs[i] = 0.0;
#pragma unroll
                    for (int j = 0; j < NA; j++)
                        if ((km >> j) & 1u) rs[i] = rs[i] + c[i][j];
                }
#pragma unroll
                for (int j = 0; j < NA; j++) {
                    cs[j] = 0.0;
#pragma unroll
                    for (int i = 0; i < NP; i++) cs[j] = cs[j] + c[i][j];
                }
                double chi2 = 0.0;
#pragma unroll
                for (int i = 0; i < NP; i++)
#pragma unroll
                    for (int j = 0; j < NA; j++)
                        if ((km >> j) & 1u) {
                            const double e = (rs[i] * cs[j]) / total;
                            const double d = c[i][j] - e;
                            chi2 += (d * d) / e;
                        }
                stat = chi2;
                const int df = NP * pk - 1;
                double cdf;
                if (chi2 != chi2)
                    cdf = nan("");
                else if (chi2 <= 0.0)
                    cdf = 0.0;
                else if (isinf(chi2))
                    cdf = 1.0;
                else
                    cdf = 1.0 - gamma_q_halfint(df, chi2 * 0.5);
                pval = 1.00 - cdf;
            } else {
                double total = 0.0;
#pragma unroll
                for (int i = 0; i < NP; i++)
#pragma unroll
                    for (int j = 0; j < NA; j++) {
                        c[i][j] = ((km >> j) & 1u) ? cnt[i][j] : 0.0;
                        if ((km >> j) & 1u) total = total + c[i][j];
                    }
                if (total > 34.0) {
                    const double coef = 34.0 / total;
#pragma unroll
                    for (int i = 0; i < NP; i++)
#pragma unroll
                        for (int j = 0; j < NA; j++)
                            if ((km >> j) & 1u) c[i][j] = floor(c[i][j] * coef);
                }
                // from here on every cell is an integer <= 34 (the rescale, fisher_exact_test.rs:51-58): the fills, the
                // marginal sums and the comparisons are exact in integer arithmetic, which leaves the FP64 pipe to the
                // log-factorial sums and 10^x
                int ci[NP][NA], ri[NP], cj[NA];
#pragma unroll
                for (int i = 0; i < NP; i++)
#pragma unroll
                    for (int j = 0; j < NA; j++) ci[i][j] = (int)fmin(c[i][j], 35.0);  // > 34 only without the rescale: never
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    ri[i] = 0;
#pragma unroll
                    for (int j = 0; j < NA; j++) ri[i] += ci[i][j];
                }
#pragma unroll
                for (int j = 0; j < NA; j++) {
                    cj[j] = 0;
#pragma unroll
                    for (int i = 0; i < NP; i++) cj[j] += ci[i][j];
                }
                double lp = 0.0;
#pragma unroll
                for (int i = 0; i < NP; i++) lp = lp + s_lf[ri[i]];
#pragma unroll
                for (int j = 0; j < NA; j++)
                    if ((km >> j) & 1u) lp = lp + s_lf[cj[j]];
                const double p_obs = ratio_i<NP, NA>(ci, km, lp, s_lf);
                const int jlast = 31 - __clz(km);  // the last kept column
                double p_ext = 0.0;
                bool panic = false;
                for (int mi = 0; mi < NP && !panic; mi++)
                    for (int mj = 0; mj < NA && !panic; mj++) {
                        if (!((km >> mj) & 1u)) continue;
                        // forward fill (fisher_exact_test.rs:78-92); `as usize` saturates negatives to 0
                        int cdown[NA];
#pragma unroll
                        for (int j = 0; j < NA; j++) cdown[j] = 0;
#pragma unroll
                        for (int i = 0; i < NP; i++) {
                            int r = 0;
#pragma unroll
                            for (int j = 0; j < NA; j++) {
                                if (!((km >> j) & 1u)) continue;
                                const int mx = min(max(ri[i] - r, 0), max(cj[j] - cdown[j], 0));
                                int v;
                                if ((i == NP - 1) | (j == jlast))
                                    v = mx;
                                else if ((i < mi) | (j < mj))
                                    v = 0;
                                else
                                    v = mx;
                                ci[i][j] = v;
                                r += v;
                                cdown[j] += v;
                            }
                        }
                        // reverse fill (fisher_exact_test.rs:96-111): the full row / column sums of the CURRENT table
                        int rfull[NP];
#pragma unroll
                        for (int i = 0; i < NP; i++) {
                            rfull[i] = 0;
#pragma unroll
                            for (int j = 0; j < NA; j++) rfull[i] += ci[i][j];
                        }
#pragma unroll
                        for (int j = NA - 1; j >= 0; j--)
#pragma unroll
                            for (int i = NP - 1; i >= 0; i--) {
                                if (!((km >> j) & 1u)) continue;
                                const int mx = min(max(ri[i] - rfull[i], 0), max(cj[j] - cdown[j], 0));
                                if (mx > 0) {
                                    const int d = mx - ci[i][j];
                                    ci[i][j] = mx;
                                    rfull[i] += d;
                                    cdown[j] += d;
                                }
                            }
#pragma unroll
                        for (int i = 0; i < NP; i++)
                            if (rfull[i] != ri[i]) panic = true;
#pragma unroll
                        for (int j = 0; j < NA; j++)
                            if (cdown[j] != cj[j]) panic = true;
                        if (!panic) p_ext += ratio_i<NP, NA>(ci, km, lp, s_lf);
                    }
                if (panic) {
                    status = PG_LOCUS_PANIC;  // assert! at fisher_exact_test.rs:113-114
                } else {
                    stat = p_obs;
                    pval = p_obs + p_ext;
                }
            }
        }
        uint64_t mv = (uint64_t)status;
        if (status == PG_LOCUS_OK) {
            mv |= (uint64_t)pk << 8;
            int a = 0;
#pragma unroll
            for (int j = 0; j < NA; j++)
                if ((km >> j) & 1u) {
                    mv |= (uint64_t)p.codes[j] << (16 + 8 * a);
                    a++;
                }
        }
        p.meta[locus] = mv;
        double *o = p.stats + (size_t)locus * 4;
        *reinterpret_cast<double2 *>(o) = make_double2(stat, nan(""));
        *reinterpret_cast<double2 *>(o + 2) = make_double2(nan(""), pval);
    }
}

template <int NP, int NA>
static cudaError_t launch_tables_t(const TableParams &p, int sm_count, cudaStream_t s) {
    const size_t smem = (size_t)kTabThreads * NA * NP * 4;
    auto kern = tables_kernel_t<NP, NA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t tiles = (p.n_loci + kTabThreads - 1) / kTabThreads;
    int64_t grid = tiles < (int64_t)sm_count * 24 ? tiles : (int64_t)sm_count * 24;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kTabThreads, smem, s>>>(p);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(kTabThreads) tables_kernel(const TableParams p) {
    extern __shared__ __align__(16) uint32_t sm_counts[];
    const int per_locus = p.A_in * p.n;  // u32 words
    const int64_t tiles = (p.n_loci + kTabThreads - 1) / kTabThreads;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t l0 = tile * kTabThreads;
        const int nl = (int)min((int64_t)kTabThreads, p.n_loci - l0);
        const size_t words = (size_t)nl * per_locus;
        const uint32_t *src = p.counts + (size_t)l0 * per_locus;
        __syncthreads();
        // the tile start is 16-byte aligned when per_locus * kTabThreads * 4 is (always: 128 loci)
        const size_t w4 = words / 4;
        for (size_t i = threadIdx.x; i < w4; i += kTabThreads)
            reinterpret_cast<uint4 *>(sm_counts)[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
        for (size_t i = w4 * 4 + threadIdx.x; i < words; i += kTabThreads) sm_counts[i] = src[i];
        __syncthreads();
        if ((int)threadIdx.x >= nl) continue;
        const int64_t locus = l0 + threadIdx.x;
        const uint32_t *cnt = sm_counts + (size_t)threadIdx.x * per_locus;
        const int n = p.n;
        int cols[PG_MAX_ALLELES];
        int pk = 0;
        int status = table_filter(cnt, p, cols, pk);
        double stat = nan(""), pval = nan("");
        if (status == PG_LOCUS_OK) {
            const int cells = n * pk;
            double c[kTabMaxCells];  // row-major n x pk
            double rs[kTabMaxPools], cs[PG_MAX_ALLELES];
            if (p.kind == PG_KIND_CHISQ) {
                // frequencies renormalised over the kept alleles (sync.rs:166-192)
                for (int i = 0; i < n; i++) {
                    double d = 0.0;
                    for (int a = 0; a < pk; a++) d = d + (double)cnt[cols[a] * n + i];
                    for (int a = 0; a < pk; a++) c[i * pk + a] = (d == 0.0) ? nan("") : (double)cnt[cols[a] * n + i] / d;
                }
                double total = 0.0;
                for (int i = 0; i < cells; i++) total = total + c[i];
                for (int i = 0; i < n; i++) {
                    rs[i] = 0.0;
                    for (int a = 0; a < pk; a++) rs[i] = rs[i] + c[i * pk + a];
                }
                for (int a = 0; a < pk; a++) {
                    cs[a] = 0.0;
                    for (int i = 0; i < n; i++) cs[a] = cs[a] + c[i * pk + a];
                }
                double chi2 = 0.0;
                for (int i = 0; i < n; i++)
                    for (int a = 0; a < pk; a++) {
                        const double e = (rs[i] * cs[a]) / total;
                        const double d = c[i * pk + a] - e;
                        chi2 += (d * d) / e;
                    }
                stat = chi2;
                const double df = (double)cells - 1.0;
                double cdf;
                if (chi2 <= 0.0)
                    cdf = 0.0;
                else if (isinf(chi2))
                    cdf = 1.0;
                else
                    cdf = gamma_lr_dev(df / 2.0, chi2 * 0.5);
                pval = 1.00 - cdf;
            } else {
                for (int i = 0; i < n; i++)
                    for (int a = 0; a < pk; a++) c[i * pk + a] = (double)cnt[cols[a] * n + i];
                double total = 0.0;
                for (int i = 0; i < cells; i++) total = total + c[i];
                if (total > 34.0) {
                    const double coef = 34.0 / total;
                    for (int i = 0; i < cells; i++) c[i] = floor(c[i] * coef);
                }
                for (int i = 0; i < n; i++) {
                    rs[i] = 0.0;
                    for (int a = 0; a < pk; a++) rs[i] = rs[i] + c[i * pk + a];
                }
                for (int a = 0; a < pk; a++) {
                    cs[a] = 0.0;
                    for (int i = 0; i < n; i++) cs[a] = cs[a] + c[i * pk + a];
                }
                double lp = 0.0;
                for (int i = 0; i < n; i++) lp = lp + c_lf10[(int)rs[i]];
                for (int a = 0; a < pk; a++) lp = lp + c_lf10[(int)cs[a]];
                const double p_obs = hypergeom_ratio_dev(c, cells, lp);
                double p_ext = 0.0;
                bool panic = false;
                for (int mi = 0; mi < n && !panic; mi++)
                    for (int mj = 0; mj < pk && !panic; mj++) {
                        for (int i = 0; i < n; i++)
                            for (int j = 0; j < pk; j++) {
                                double r = 0.0, s = 0.0;
                                for (int jj = 0; jj < j; jj++) r = r + c[i * pk + jj];
                                for (int ii = 0; ii < i; ii++) s = s + c[ii * pk + j];
                                const double a = as_usize_f64(rs[i] - r), b = as_usize_f64(cs[j] - s);
                                const double mx = a < b ? a : b;
                                if ((i == n - 1) | (j == pk - 1))
                                    c[i * pk + j] = mx;
                                else if ((i < mi) | (j < mj))
                                    c[i * pk + j] = 0.0;
                                else
                                    c[i * pk + j] = mx;
                            }
                        for (int ij = 0; ij < pk; ij++)
                            for (int ii2 = 0; ii2 < n; ii2++) {
                                const int j = pk - (ij + 1), i = n - (ii2 + 1);
                                double r = 0.0, s = 0.0;
                                for (int jj = 0; jj < pk; jj++) r = r + c[i * pk + jj];
                                for (int ii = 0; ii < n; ii++) s = s + c[ii * pk + j];
                                const double a = as_usize_f64(rs[i] - r), b = as_usize_f64(cs[j] - s);
                                const double mx = a < b ? a : b;
                                if (mx > 0.0) c[i * pk + j] = mx;
                            }
                        for (int i = 0; i < n; i++) {
                            double r = 0.0;
                            for (int j = 0; j < pk; j++) r = r + c[i * pk + j];
                            if (r != rs[i]) panic = true;
                        }
                        for (int j = 0; j < pk; j++) {
                            double s = 0.0;
                            for (int i = 0; i < n; i++) s = s + c[i * pk + j];
                            if (s != cs[j]) panic = true;
                        }
                        if (!panic) p_ext += hypergeom_ratio_dev(c, cells, lp);
                    }
                if (panic) {
                    status = PG_LOCUS_PANIC;  // assert! at fisher_exact_test.rs:113-114
                } else {
                    stat = p_obs;
                    pval = p_obs + p_ext;
                }
            }
        }
        uint64_t mv = (uint64_t)status;
        if (status == PG_LOCUS_OK) {
            mv |= (uint64_t)pk << 8;
            for (int a = 0; a < pk; a++) mv |= (uint64_t)p.codes[cols[a]] << (16 + 8 * a);
        }
        p.meta[locus] = mv;
        double *o = p.stats + (size_t)locus * 4;
        *reinterpret_cast<double2 *>(o) = make_double2(stat, nan(""));
        *reinterpret_cast<double2 *>(o + 2) = make_double2(nan(""), pval);
    }
}

// ---- more than 16 pools: one warp per locus, lane = pool (stride 32), counts re-read from L1/L2 ---------------------
// LocusCounts::filter by the whole warp.  The keep-mask is the reference's: a pooled frequency within the rounding
// bound of a threshold is re-evaluated by one lane in pool order with separately rounded multiply and add
// (src/base/sync.rs:258-271).  Every lane returns the same status, kept columns cols[0..pk).
__device__ __forceinline__ int wide_filter(const TableParams &p, const uint32_t *cnt, const int *col, int pc, int lane,
                                           int *cols, int &pk) {
    const int n = p.n;
    const double tol_rel = 2.0 * ((double)n + 8.0) * kEps;
    double q[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
    double dmin = 1.7976931348623157e308;
    int miss = 0;
    for (int i = lane; i < n; i += 32) {
        double d = 0.0;
        for (int a = 0; a < pc; a++) d += (double)cnt[col[a] * n + i];
        dmin = fmin(dmin, d);
        if (d == 0.0) {
            miss++;
        } else {
            const double wi = p.w[i];
            for (int a = 0; a < pc; a++) q[a] += ((double)cnt[col[a] * n + i] / d) * wi;
        }
    }
    for (int off = 16; off >= 1; off >>= 1) {
        dmin = fmin(dmin, __shfl_xor_sync(PG_FULL_MASK, dmin, off));
        miss += __shfl_xor_sync(PG_FULL_MASK, miss, off);
        for (int a = 0; a < pc; a++) q[a] += __shfl_xor_sync(PG_FULL_MASK, q[a], off);
    }
    pk = 0;
    if (dmin < p.min_depth_f) return PG_LOCUS_FILTERED;
    for (int a = 0; a < pc; a++) {
        double qa = q[a];
        const double tl = tol_rel * fmax(fabs(qa), 1.0);
        if (fabs(qa - p.maf) <= tl || fabs(qa - p.one_minus_maf) <= tl) {
            if (lane == 0) {  // the reference's order decides
                double e = 0.0;
                for (int i = 0; i < n; i++) {
                    double d = 0.0;
                    for (int b = 0; b < pc; b++) d += (double)cnt[col[b] * n + i];
                    const double f = (d == 0.0) ? nan("") : (double)cnt[col[a] * n + i] / d;
                    const double term = (f != f) ? 0.0 : __dmul_rn(f, p.w[i]);
                    e = __dadd_rn(e, term);
                }
                qa = e;
            }
            qa = __shfl_sync(PG_FULL_MASK, qa, 0);
        }
        if (!((qa < p.maf) | (qa > p.one_minus_maf))) cols[pk++] = col[a];
    }
    if (pk < 2) return PG_LOCUS_FILTERED;
    if (miss == n) return PG_LOCUS_FILTERED;
    if (((double)miss / (double)n) > p.max_miss) return PG_LOCUS_FILTERED;
    return PG_LOCUS_OK;
}

// tables::chisq: three passes (depth + pooled frequencies, column sums, chi-square)
__global__ void __launch_bounds__(256) chisq_wide_kernel(const TableParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = p.n;
    int col[PG_MAX_ALLELES];
    int pc = 0;
    for (int j = 0; j < p.A_in; j++)
        if (j != p.drop_col) col[pc++] = j;
    for (int64_t locus = gw; locus < p.n_loci; locus += nw) {
        const uint32_t *cnt = p.counts + (size_t)locus * p.A_in * n;
        int cols[PG_MAX_ALLELES];
        int pk = 0;
        const int status = wide_filter(p, cnt, col, pc, lane, cols, pk);
        double stat = nan(""), pval = nan("");
        if (status == PG_LOCUS_OK) {
            // pass 2: column sums and the total of the frequencies renormalised over the kept alleles (sync.rs:166-192)
            double cs[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
            double total = 0.0;
            for (int i = lane; i < n; i += 32) {
                double d = 0.0;
                for (int a = 0; a < pk; a++) d = d + (double)cnt[cols[a] * n + i];
                for (int a = 0; a < pk; a++) {
                    const double c = (d == 0.0) ? nan("") : (double)cnt[cols[a] * n + i] / d;
                    cs[a] += c;
                    total += c;
                }
            }
            for (int off = 16; off >= 1; off >>= 1) {
                total += __shfl_xor_sync(PG_FULL_MASK, total, off);
                for (int a = 0; a < pk; a++) cs[a] += __shfl_xor_sync(PG_FULL_MASK, cs[a], off);
            }
            // pass 3: chi-square (chisq_test.rs:17-31)
            double chi2 = 0.0;
            for (int i = lane; i < n; i += 32) {
                double d = 0.0;
                for (int a = 0; a < pk; a++) d = d + (double)cnt[cols[a] * n + i];
                double c[PG_MAX_ALLELES], rs = 0.0;
                for (int a = 0; a < pk; a++) {
                    c[a] = (d == 0.0) ? nan("") : (double)cnt[cols[a] * n + i] / d;
                    rs = rs + c[a];
                }
                for (int a = 0; a < pk; a++) {
                    const double e = (rs * cs[a]) / total;
                    const double dd = c[a] - e;
                    chi2 += (dd * dd) / e;
                }
            }
            for (int off = 16; off >= 1; off >>= 1) chi2 += __shfl_xor_sync(PG_FULL_MASK, chi2, off);
            stat = chi2;
            const double df = (double)(n * pk) - 1.0;
            double cdf;
            if (chi2 <= 0.0)
                cdf = 0.0;
            else if (isinf(chi2))
                cdf = 1.0;
            else
                cdf = gamma_lr_dev(df / 2.0, chi2 * 0.5);
            pval = 1.00 - cdf;
        }
        if (lane == 0) {
            uint64_t mv = (uint64_t)status;
            if (status == PG_LOCUS_OK) {
                mv |= (uint64_t)pk << 8;
                for (int a = 0; a < pk; a++) mv |= (uint64_t)p.codes[cols[a]] << (16 + 8 * a);
            }
            p.meta[locus] = mv;
            double *o = p.stats + (size_t)locus * 4;
            *reinterpret_cast<double2 *>(o) = make_double2(stat, nan(""));
            *reinterpret_cast<double2 *>(o + 2) = make_double2(nan(""), pval);
        }
    }
}

// tables::fisher (src/tables/fisher_exact_test.rs:32-130) for any number of pools.  The table is rescaled to a total of
// at most 34 (51-58), so whatever the pool count at most 34 rows are non-zero; a zero row stays zero in every fill
// (its cells are bounded by its row sum) and adds log10(0!) = 0 to every sum, so only the non-zero rows are kept, with
// their original index for the two places the enumeration looks at it: `i == n - 1` and `i < max_i` (the latter only
// through the NUMBER of non-zero rows above max_i -- its "class").  All cell arithmetic is on integers <= 34, hence
// exact in any form; the floating-point sums (log10-factorials in row-major order, p_extremes over (max_i, max_j) in
// loop order) keep the reference's order.  The forward fill (78-92) overwrites every cell from cells it has already
// rewritten, so the (class, max_j) tables are independent: the lanes of the warp take them in parallel, each in its own
// scratch table, and lane 0 adds the ratios in the reference's loop order.
constexpr int kFwRows = 34;
constexpr int kFwWarps = 4;
constexpr int kFwPitch = 8;  // bytes per scratch row: 6 cells + the running row sum

__global__ void __launch_bounds__(kFwWarps * 32) fisher_wide_kernel(const TableParams p) {
    __shared__ uint8_t s_obs[kFwWarps][kFwRows * kFwPitch];   // observed (rescaled) non-zero rows | row sum in [6]
    __shared__ int s_orow[kFwWarps][kFwRows];                 // original row index
    __shared__ uint8_t s_tab[kFwWarps][32][kFwRows * kFwPitch];
    __shared__ double s_ratio[kFwWarps][(kFwRows + 1) * PG_MAX_ALLELES];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * kFwWarps + wib;
    const int64_t nw = (int64_t)gridDim.x * kFwWarps;
    const int n = p.n;
    int col[PG_MAX_ALLELES];
    int pc = 0;
    for (int j = 0; j < p.A_in; j++)
        if (j != p.drop_col) col[pc++] = j;
    uint8_t *obs = s_obs[wib];
    int *orow = s_orow[wib];
    uint8_t *T = s_tab[wib][lane];
    double *ratio = s_ratio[wib];
    for (int64_t locus = gw; locus < p.n_loci; locus += nw) {
        const uint32_t *cnt = p.counts + (size_t)locus * p.A_in * n;
        int cols[PG_MAX_ALLELES];
        int pk = 0;
        int status = wide_filter(p, cnt, col, pc, lane, cols, pk);
        double stat = nan(""), pval = nan("");
        if (status == PG_LOCUS_OK) {
            // the table's total (integers: exact in any order), the rescale to <= 34 (fisher_exact_test.rs:50-58)
            double total = 0.0;
            for (int i = lane; i < n; i += 32)
                for (int a = 0; a < pk; a++) total += (double)cnt[cols[a] * n + i];
            for (int off = 16; off >= 1; off >>= 1) total += __shfl_xor_sync(PG_FULL_MASK, total, off);
            const bool scale = total > 34.0;
            const double coef = 34.0 / total;
            int nz = 0;
            int cs[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
            __syncwarp();
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                int cell[PG_MAX_ALLELES], rsum = 0;
                for (int a = 0; a < pk; a++) {
                    double c = (i < n) ? (double)cnt[cols[a] * n + i] : 0.0;
                    if (scale) c = floor(c * coef);
                    cell[a] = (int)c;
                    rsum += cell[a];
                    cs[a] += cell[a];
                }
                const unsigned nzm = __ballot_sync(PG_FULL_MASK, rsum > 0);
                if (rsum > 0) {
                    const int pos = nz + __popc(nzm & ((1u << lane) - 1u));
                    if (pos < kFwRows) {  // the total is <= 34, so at most 34 rows carry a count
                        for (int a = 0; a < pk; a++) obs[pos * kFwPitch + a] = (uint8_t)cell[a];
                        obs[pos * kFwPitch + 6] = (uint8_t)rsum;
                        orow[pos] = i;
                    }
                }
                nz += __popc(nzm);
            }
            for (int a = 0; a < pk; a++) cs[a] = __reduce_add_sync(PG_FULL_MASK, cs[a]);
            if (nz > kFwRows) nz = kFwRows;
            __syncwarp();
            // log-product of the marginal sums (60-67) and the observed ratio (69), sequential like the reference
            double lp = 0.0;
            int tot2 = 0;
            for (int r = 0; r < nz; r++) {
                lp = lp + c_lf10[obs[r * kFwPitch + 6]];
                tot2 += obs[r * kFwPitch + 6];
            }
            for (int a = 0; a < pk; a++) lp = lp + c_lf10[cs[a]];
            double so = 0.0;
            for (int r = 0; r < nz; r++)
                for (int a = 0; a < pk; a++) so = so + c_lf10[obs[r * kFwPitch + a]];
            so = so + c_lf10[tot2];
            const double p_obs = pow(10.0, lp - so);
            // the (class, max_j) tables, one per lane at a time
            const int ntask = (nz + 1) * pk;
            bool panic = false;
            for (int t = lane; t < ntask; t += 32) {
                const int cls = t / pk, mj = t - cls * pk;
                int S[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
                for (int r = 0; r < nz; r++) {  // forward fill (78-92)
                    const int rsr = obs[r * kFwPitch + 6];
                    const bool last_row = orow[r] == n - 1;
                    int rr = 0;
                    for (int j = 0; j < pk; j++) {
                        const int a = max(rsr - rr, 0), b = max(cs[j] - S[j], 0);
                        const int mx = min(a, b);
                        const int v = (last_row | (j == pk - 1)) ? mx : (((r < cls) | (j < mj)) ? 0 : mx);
                        T[r * kFwPitch + j] = (uint8_t)v;
                        rr += v;
                        S[j] += v;
                    }
                    T[r * kFwPitch + 6] = (uint8_t)rr;
                }
                for (int j = pk - 1; j >= 0; j--)  // reverse fill (96-111)
                    for (int r = nz - 1; r >= 0; r--) {
                        const int R = T[r * kFwPitch + 6];
                        const int a = max((int)obs[r * kFwPitch + 6] - R, 0), b = max(cs[j] - S[j], 0);
                        const int mx = min(a, b);
                        if (mx > 0) {
                            const int old = T[r * kFwPitch + j];
                            T[r * kFwPitch + j] = (uint8_t)mx;
                            T[r * kFwPitch + 6] = (uint8_t)(R + mx - old);
                            S[j] += mx - old;
                        }
                    }
                int tsum = 0;
                for (int r = 0; r < nz; r++) {  // the marginal sums must have been kept (113-114)
                    if (T[r * kFwPitch + 6] != obs[r * kFwPitch + 6]) panic = true;
                    tsum += T[r * kFwPitch + 6];
                }
                for (int j = 0; j < pk; j++)
                    if (S[j] != cs[j]) panic = true;
                double sr = 0.0;
                for (int r = 0; r < nz; r++)
                    for (int j = 0; j < pk; j++) sr = sr + c_lf10[T[r * kFwPitch + j]];
                sr = sr + c_lf10[min(tsum, 35)];
                ratio[t] = pow(10.0, lp - sr);
            }
            panic = __any_sync(PG_FULL_MASK, panic);
            __syncwarp();
            if (panic) {
                status = PG_LOCUS_PANIC;  // assert! at fisher_exact_test.rs:113-114
            } else {
                double p_ext = 0.0;
                if (lane == 0) {  // p_extremes in the reference's loop order (max_i outer, max_j inner)
                    int cls = 0;
                    for (int mi = 0; mi < n; mi++) {
                        while (cls < nz && orow[cls] < mi) cls++;
                        for (int mj = 0; mj < pk; mj++) p_ext += ratio[cls * pk + mj];
                    }
                }
                stat = p_obs;
                pval = p_obs + p_ext;
            }
            __syncwarp();
        }
        if (lane == 0) {
            uint64_t mv = (uint64_t)status;
            if (status == PG_LOCUS_OK) {
                mv |= (uint64_t)pk << 8;
                for (int a = 0; a < pk; a++) mv |= (uint64_t)p.codes[cols[a]] << (16 + 8 * a);
            }
            p.meta[locus] = mv;
            double *o = p.stats + (size_t)locus * 4;
            *reinterpret_cast<double2 *>(o) = make_double2(stat, nan(""));
            *reinterpret_cast<double2 *>(o + 2) = make_double2(nan(""), pval);
        }
    }
}

cudaError_t launch_tables(const TableParams &p, int sm_count, cudaStream_t s) {
    static bool lf_ready[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!lf_ready[dev & 63]) {
        double lf[36];
        for (int x = 0; x < 36; x++) {
            double out = 0.0;
            for (int i = 2; i < x + 1; i++) out = out + log10((double)i);
            lf[x] = out;
        }
        // on the launch stream (pageable source: staged before the call returns), so the kernel below is ordered after it
        cudaError_t e = cudaMemcpyToSymbolAsync(c_lf10, lf, sizeof lf, 0, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
        lf_ready[dev & 63] = true;
    }
    static const bool force_wide = getenv("PG_TABLES_WIDE") != nullptr;  // tests: the wide kernels on small tables too
    if (p.n > kTabMaxPools || force_wide) {
        if (p.kind == PG_KIND_CHISQ) {
            int64_t grid = (p.n_loci * 32 + 255) / 256;
            if (grid > (int64_t)sm_count * 16) grid = (int64_t)sm_count * 16;
            if (grid < 1) grid = 1;
            chisq_wide_kernel<<<(unsigned)grid, 256, 0, s>>>(p);
        } else {
            int64_t grid = (p.n_loci + kFwWarps - 1) / kFwWarps;
            if (grid > (int64_t)sm_count * 4) grid = (int64_t)sm_count * 4;
            if (grid < 1) grid = 1;
            fisher_wide_kernel<<<(unsigned)grid, kFwWarps * 32, 0, s>>>(p);
        }
        return cudaGetLastError();
    }
    // register-resident kernels for the common small tables; everything else takes the generic kernel
    if (p.n == 2) {
        if (p.A_in == 4) return launch_tables_t<2, 4>(p, sm_count, s);
        if (p.A_in == 5) return launch_tables_t<2, 5>(p, sm_count, s);
        if (p.A_in == 6) return launch_tables_t<2, 6>(p, sm_count, s);
    } else if (p.n == 3) {
        if (p.A_in == 4) return launch_tables_t<3, 4>(p, sm_count, s);
        if (p.A_in == 6) return launch_tables_t<3, 6>(p, sm_count, s);
    } else if (p.n == 4) {
        if (p.A_in == 4) return launch_tables_t<4, 4>(p, sm_count, s);
        if (p.A_in == 6) return launch_tables_t<4, 6>(p, sm_count, s);
    }
    const size_t smem = (size_t)kTabThreads * p.A_in * p.n * 4;
    cudaError_t e = cudaFuncSetAttribute(tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t tiles = (p.n_loci + kTabThreads - 1) / kTabThreads;
    int64_t grid = tiles < (int64_t)sm_count * 8 ? tiles : (int64_t)sm_count * 8;
    if (grid < 1) grid = 1;
    tables_kernel<<<(unsigned)grid, kTabThreads, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace pg
