// pg_scan.cuh -- the per-locus OLS / Pearson scan kernel (ols_iterate: src/gwas/ols.rs:201-276,
// correlation: src/gwas/correlation_test.rs:73-129) for sm_100a.
//
// One persistent CTA per SM.  Every warp owns a private ring of kNBuf shared-memory buffers fed by
// 1-D bulk copies (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), so there is no CTA-wide
// synchronisation in the steady state.  A warp works on groups of G consecutive loci:
//   phase 1 (per locus)  stream the [A][rows] frequency chunks, lane = pool, accumulate the column
//                        sums, the A(A+1)/2 products and the A*K cross products with the centred
//                        phenotypes (kept in shared memory), reduce 32 accumulators with 31 shuffles,
//                        then evaluate the reference's keep-mask (LocusCounts::filter,
//                        src/base/sync.rs:195-303) and allele order (src/base/sync.rs:478-505).
//                        Decisions that land within a rounding bound of a threshold are re-evaluated
//                        in the reference's exact sequential order (bit-exact mask).
//   phase 2 (per group)  lane = locus: centred normal equations, Cholesky, beta, residual variance;
//   phase 3 (per group)  lane = (locus, allele, phenotype): t and the Student-t two-sided p-value.
#pragma once
#include "pg_device.cuh"
#include "pg_internal.h"

namespace pg {

constexpr int kNBuf = 2;
constexpr int kRedPitchDecl = 33;
constexpr int kBufRows = kChunkRows;

template <int A, int K, bool W>
struct Acc {
    static constexpr int S0 = 0;
    static constexpr int P0 = S0 + A;
    static constexpr int C0 = P0 + A * (A + 1) / 2;
    static constexpr int Q0 = C0 + A * K;
    static constexpr int N = Q0 + (W ? A : 0);
    static constexpr int NB = (N + 31) / 32;
    static constexpr int NP = NB * 32;
    __host__ __device__ static constexpr int tri(int j, int l) {  // j <= l
        return P0 + j * A - j * (j - 1) / 2 + (l - j);
    }
};

// loci per warp group: fill the 32 lanes of phase 3 as well as possible
__host__ __device__ constexpr int group_size(int T) {
    int best = 1;
    int best_num = 0, best_den = 1;
    for (int g = 1; g <= 16; g++) {
        const int tasks = g * T;
        const int passes = (tasks + 31) / 32;
        // efficiency tasks / (32 * passes); the smaller g wins ties (less shared memory per warp)
        if ((long long)tasks * best_den > (long long)best_num * (32 * passes)) {
            best = g;
            best_num = tasks;
            best_den = 32 * passes;
        }
    }
    return best;
}

template <int A, int K, bool W>
struct WarpSmem {
    using AC = Acc<A, K, W>;
    static constexpr int T = (A - 1) * K;
    static constexpr int G = group_size(T);
    static constexpr int bar_bytes = 32;
    static constexpr int fbuf_bytes = A * kBufRows * 8;
    static constexpr int dbuf_bytes = kBufRows * 4;
    static constexpr int tot_bytes = G * AC::NP * 8;
    static constexpr int sel_bytes = ((G * 8 + 15) / 16) * 16;
    static constexpr int tb_bytes = G * T * 16;
    // the chunk buffer that was consumed last doubles as the reduction scratch: rows of 33 doubles
    static constexpr int RED_ROWS_CAP = fbuf_bytes / (kRedPitchDecl * 8);
    static constexpr int RED_ROWS = RED_ROWS_CAP < 32 ? RED_ROWS_CAP : 32;
    static constexpr int off_f = bar_bytes;
    static constexpr int off_d = off_f + kNBuf * fbuf_bytes;
    static constexpr int off_tot = off_d + kNBuf * dbuf_bytes;
    static constexpr int off_sel = off_tot + tot_bytes;
    static constexpr int off_tb = off_sel + sel_bytes;
    static constexpr int bytes = ((off_tb + tb_bytes + 127) / 128) * 128;
};

__host__ __device__ inline size_t scan_common_bytes(int K, int n_pad, bool weighted) {
    size_t b = (size_t)(K + (weighted ? 1 : 0)) * n_pad * 8;
    return (b + 127) / 128 * 128;
}

template <int A, int K, bool W, bool NANAWARE>
__device__ __forceinline__ void accum_row(double (&acc)[Acc<A, K, W>::NP], const double (&f)[A],
                                          const double (&y)[K], double w) {
    using AC = Acc<A, K, W>;
#pragma unroll
    for (int j = 0; j < A; j++) {
        if (NANAWARE)
            acc[AC::S0 + j] += (f[j] != f[j]) ? 0.0 : f[j];  // column sums ignore NaN (sync.rs:483-489)
        else
            acc[AC::S0 + j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < A; j++)
#pragma unroll
        for (int l = j; l < A; l++) acc[AC::tri(j, l)] = fma(f[j], f[l], acc[AC::tri(j, l)]);
#pragma unroll
    for (int j = 0; j < A; j++)
#pragma unroll
        for (int k = 0; k < K; k++) acc[AC::C0 + j * K + k] = fma(f[j], y[k], acc[AC::C0 + j * K + k]);
    if (W) {
#pragma unroll
        for (int j = 0; j < A; j++) acc[AC::Q0 + j] = fma(f[j], w, acc[AC::Q0 + j]);
    }
}

// Warp reduction of the N accumulators through shared memory: every lane stores its partial sums into a
// [accumulator][lane] tile (row pitch 33 doubles, conflict free both ways), then lane a adds up row a.  ~3x fewer
// instructions than a shuffle tree and it keeps the accumulators in aligned 64-bit register pairs.
constexpr int kRedPitch = 33;
template <int N, int NP, int R>
__device__ __forceinline__ void reduce_store(const double (&acc)[NP], double *red, double *tot, int lane) {
#pragma unroll
    for (int b = 0; b < (N + R - 1) / R; b++) {
        if (b) __syncwarp();
#pragma unroll
        for (int i = 0; i < R; i++)
            if (b * R + i < N) red[i * kRedPitch + lane] = acc[b * R + i];
        __syncwarp();
        if (lane < R && b * R + lane < N) {
            const double *row = red + lane * kRedPitch;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                s0 += row[i];
                s1 += row[i + 1];
                s2 += row[i + 2];
                s3 += row[i + 3];
            }
            tot[b * R + lane] = (s0 + s1) + (s2 + s3);
        }
    }
}

// q_j = sum_i f_ij * (s_i / S), accumulated in pool order with separately rounded multiply and add,
// NaN frequencies contribute 0 -- the exact arithmetic of src/base/sync.rs:258-271.
static __device__ __noinline__ double exact_q(const ScanParams &p, int64_t locus, int j, const double *ws) {
    const Layout &lay = p.lay;
    const double *fl = p.freq + (size_t)locus * lay.freq_stride();
    double q = 0.0;
    int i0 = 0;
    for (int c = 0; c < lay.n_chunks; c++) {
        const int rcc = (c == lay.n_chunks - 1) ? lay.rc_last : lay.rc;
        const double *col = fl + (size_t)c * lay.A * lay.rc + (size_t)j * rcc;
        const int rows = min(rcc, lay.n - i0);
        for (int r = 0; r < rows; r++) {
            const double f = col[r];
            const double w = ws ? ws[i0 + r] : p.w_uniform;
            const double term = (f != f) ? 0.0 : __dmul_rn(f, w);
            q = __dadd_rn(q, term);
        }
        i0 += rcc;
    }
    return q;
}

// renormalised frequency of (pool i, column j) over the kept columns:
// c_ij / sum_{kept} c_i.  (LocusCounts::to_frequencies after the filter, src/base/sync.rs:166-192),
// with c_ij = rint(f_ij * depth_i) recovering the integer count exactly.
template <int A>
__device__ __forceinline__ void renorm_row(const double (&f)[A], uint32_t d, unsigned kept, double (&F)[A]) {
    double c[A];
    double dk = 0.0;
#pragma unroll
    for (int j = 0; j < A; j++) {
        c[j] = (d == 0u) ? 0.0 : rint(f[j] * (double)d);
        if ((kept >> j) & 1u) dk += c[j];
    }
#pragma unroll
    for (int j = 0; j < A; j++) {
        if ((kept >> j) & 1u)
            F[j] = (dk == 0.0) ? nan("") : c[j] / dk;
        else
            F[j] = 0.0;
    }
}

// column sum of the renormalised frequencies in pool order, NaN ignored (src/base/sync.rs:483-489)
template <int A>
__device__ __noinline__ double exact_colsum(const ScanParams &p, int64_t locus, int jsel, unsigned kept) {
    const Layout &lay = p.lay;
    const double *fl = p.freq + (size_t)locus * lay.freq_stride();
    const uint32_t *dl = p.depth + (size_t)locus * lay.depth_stride();
    double s = 0.0;
    int i0 = 0;
    for (int c = 0; c < lay.n_chunks; c++) {
        const int rcc = (c == lay.n_chunks - 1) ? lay.rc_last : lay.rc;
        const double *blk = fl + (size_t)c * lay.A * lay.rc;
        const int rows = min(rcc, lay.n - i0);
        for (int r = 0; r < rows; r++) {
            double f[A], F[A];
#pragma unroll
            for (int j = 0; j < A; j++) f[j] = blk[(size_t)j * rcc + r];
            renorm_row<A>(f, dl[i0 + r], kept, F);
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < A; j++)
                if (j == jsel) v = F[j];
            if (v == v) s = __dadd_rn(s, v);
        }
        i0 += rcc;
    }
    return s;
}

// Full reference pipeline for one locus straight from global memory: exact q with NaN handling,
// missingness, renormalisation over the kept alleles.  Used when a pool has no coverage or when a
// removed allele carries reads (both rare).  Leaves the re-accumulated totals in tot[].
template <int A, int K, bool W>
__device__ __noinline__ int slow_locus(const ScanParams &p, int64_t locus, const double *ys, const double *ws,
                                       double *red, double *tot, int lane, bool decide, unsigned &kept_out) {
    using AC = Acc<A, K, W>;
    const Layout &lay = p.lay;
    const double *fl = p.freq + (size_t)locus * lay.freq_stride();
    const uint32_t *dl = p.depth + (size_t)locus * lay.depth_stride();
    // 1. keep-mask of the alleles (lane j evaluates column j sequentially)
    unsigned kept = kept_out;
    if (decide) {  // a pool without coverage poisoned the fast sums: redo the MAF decision exactly
        bool keep_j = false;
        if (lane < A) {
            const double q = exact_q(p, locus, lane, ws);
            keep_j = !((q < p.maf) | (q > p.one_minus_maf));
        }
        kept = __ballot_sync(PG_FULL_MASK, keep_j) & ((1u << A) - 1u);
        kept_out = kept;
        if (__popc(kept) < 2) return PG_LOCUS_FILTERED;
    }
    // 2. missingness on the first kept column: NaN <=> the pool has no coverage (sync.rs:287-299)
    if (decide) {
        int miss = 0;
        for (int i = lane; i < lay.n; i += 32) miss += (dl[i] == 0u) ? 1 : 0;
        miss = __reduce_add_sync(PG_FULL_MASK, miss);
        if (miss == lay.n) return PG_LOCUS_FILTERED;
        if (((double)miss / (double)lay.n) > p.max_miss) return PG_LOCUS_FILTERED;
    }
    // 3. re-accumulate over the renormalised frequencies
    double acc[AC::NP];
#pragma unroll
    for (int i = 0; i < AC::NP; i++) acc[i] = 0.0;
    int i0 = 0;
    for (int c = 0; c < lay.n_chunks; c++) {
        const int rcc = (c == lay.n_chunks - 1) ? lay.rc_last : lay.rc;
        const double *blk = fl + (size_t)c * lay.A * lay.rc;
        const int rows = min(rcc, lay.n - i0);
        for (int r = lane; r < rows; r += 32) {
            double f[A], F[A], y[K];
#pragma unroll
            for (int j = 0; j < A; j++) f[j] = blk[(size_t)j * rcc + r];
            renorm_row<A>(f, dl[i0 + r], kept, F);
#pragma unroll
            for (int k = 0; k < K; k++) y[k] = ys[k * lay.n_pad + i0 + r];
            accum_row<A, K, W, true>(acc, F, y, 0.0);
        }
        i0 += rcc;
    }
    __syncwarp();
    reduce_store<AC::N, AC::NP, WarpSmem<A, K, W>::RED_ROWS>(acc, red, tot, lane);
    __syncwarp();
    return PG_LOCUS_OK;
}

__device__ __forceinline__ uint64_t pack_sel(int status, int nslots, const int *cols, unsigned kept) {
    uint64_t v = (uint64_t)(status & 0xff) | ((uint64_t)(nslots & 0xff) << 8) | ((uint64_t)(kept & 0x3f) << 40);
    for (int s = 0; s < PG_MAX_SLOTS; s++) v |= (uint64_t)(cols[s] & 0xf) << (16 + 4 * s);
    return v;
}

// Cholesky of the m x m SPD matrix S (lower triangle filled) -> Li = L^-1 (lower), dg = diag(S^-1)
__device__ __forceinline__ bool chol_inv(int m, const double (&S)[PG_MAX_SLOTS][PG_MAX_SLOTS],
                                         double (&Li)[PG_MAX_SLOTS][PG_MAX_SLOTS], double (&dg)[PG_MAX_SLOTS]) {
    double Lm[PG_MAX_SLOTS][PG_MAX_SLOTS];
    bool ok = true;
    for (int a = 0; a < m; a++)
        for (int b = 0; b <= a; b++) {
            double s = S[a][b];
            for (int c = 0; c < b; c++) s -= Lm[a][c] * Lm[b][c];
            if (a == b) {
                if (!(s > 0.0)) {
                    ok = false;
                    s = 1.0;
                }
                Lm[a][a] = sqrt(s);
            } else {
                Lm[a][b] = s / Lm[b][b];
            }
        }
    for (int a = 0; a < m; a++) {
        Li[a][a] = 1.0 / Lm[a][a];
        for (int b = 0; b < a; b++) {
            double s = 0.0;
            for (int c = b; c < a; c++) s -= Lm[a][c] * Li[c][b];
            Li[a][b] = s / Lm[a][a];
        }
    }
    for (int b = 0; b < m; b++) {
        double s = 0.0;
        for (int a = b; a < m; a++) s += Li[a][b] * Li[a][b];
        dg[b] = s;
    }
    return ok;
}

// beta_k = S^-1 sxy_k through Li; returns z'z (= beta' sxy)
__device__ __forceinline__ double solve_rhs(int m, const double (&Li)[PG_MAX_SLOTS][PG_MAX_SLOTS],
                                            const double (&sxy)[PG_MAX_SLOTS], double (&beta)[PG_MAX_SLOTS]) {
    double z[PG_MAX_SLOTS];
    double zz = 0.0;
    for (int a = 0; a < m; a++) {
        double s = 0.0;
        for (int b = 0; b <= a; b++) s += Li[a][b] * sxy[b];
        z[a] = s;
        zz += s * s;
    }
    for (int b = 0; b < m; b++) {
        double s = 0.0;
        for (int a = b; a < m; a++) s += Li[a][b] * z[a];
        beta[b] = s;
    }
    return zz;
}

// Reference-style two-pass evaluation of one locus straight from global memory (explicit centred moments, explicit
// residuals e = y - Xb as src/gwas/ols.rs:98-104): used when the single-pass Gram form would lose digits to
// cancellation (near-perfect fits, nearly constant or nearly collinear allele columns).  One lane per locus.
template <int A, int K>
__device__ __noinline__ int explicit_locus(const ScanParams &p, int64_t locus, unsigned kept, int m, const int *cols,
                                           const double *xbar, const double *ys, double *tbg) {
    const Layout &lay = p.lay;
    const double *fl = p.freq + (size_t)locus * lay.freq_stride();
    const uint32_t *dl = p.depth + (size_t)locus * lay.depth_stride();
    const double nn = (double)lay.n;
    double S[PG_MAX_SLOTS][PG_MAX_SLOTS], sxy[K][PG_MAX_SLOTS], ybar[K];
    for (int a = 0; a < PG_MAX_SLOTS; a++) {
        for (int b = 0; b < PG_MAX_SLOTS; b++) S[a][b] = 0.0;
        for (int k = 0; k < K; k++) sxy[k][a] = 0.0;
    }
    for (int k = 0; k < K; k++) ybar[k] = p.ysum[k] / nn;
    for (int pass = 0; pass < 2; pass++) {
        double Li[PG_MAX_SLOTS][PG_MAX_SLOTS], dg[PG_MAX_SLOTS], beta[K][PG_MAX_SLOTS], rss[K];
        if (pass == 1) {
            if (p.kind == PG_KIND_OLS) {
                if (!chol_inv(m, S, Li, dg)) return PG_LOCUS_FAILED;
                for (int k = 0; k < K; k++) {
                    solve_rhs(m, Li, sxy[k], beta[k]);
                    rss[k] = 0.0;
                }
            } else {
                for (int a = 0; a < m; a++)
                    for (int k = 0; k < K; k++) {
                        tbg[(a * K + k) * 2 + 0] = sxy[k][a] / (sqrt(S[a][a]) * sqrt(p.syy[k]));
                        tbg[(a * K + k) * 2 + 1] = 0.0;
                    }
                return PG_LOCUS_OK;
            }
        }
        int i0 = 0;
        for (int c = 0; c < lay.n_chunks; c++) {
            const int rcc = (c == lay.n_chunks - 1) ? lay.rc_last : lay.rc;
            const double *blk = fl + (size_t)c * lay.A * lay.rc;
            const int rows = min(rcc, lay.n - i0);
            for (int r = 0; r < rows; r++) {
                double f[A], F[A], x[PG_MAX_SLOTS];
#pragma unroll
                for (int j = 0; j < A; j++) f[j] = blk[(size_t)j * rcc + r];
                renorm_row<A>(f, dl[i0 + r], kept, F);
                for (int a = 0; a < m; a++) {
                    double v = 0.0;
#pragma unroll
                    for (int j = 0; j < A; j++)
                        if (j == cols[a]) v = F[j];
                    x[a] = v - xbar[a];
                }
                if (pass == 0) {
                    for (int a = 0; a < m; a++) {
                        for (int b = 0; b <= a; b++) S[a][b] = fma(x[a], x[b], S[a][b]);
                        for (int k = 0; k < K; k++) sxy[k][a] = fma(x[a], ys[k * lay.n_pad + i0 + r] - ybar[k], sxy[k][a]);
                    }
                } else {
                    for (int k = 0; k < K; k++) {
                        double e = ys[k * lay.n_pad + i0 + r] - ybar[k];
                        for (int a = 0; a < m; a++) e -= beta[k][a] * x[a];
                        rss[k] = fma(e, e, rss[k]);
                    }
                }
            }
            i0 += rcc;
        }
        if (pass == 1) {
            const double dfe = nn - (double)(m + 1);
            for (int k = 0; k < K; k++) {
                const double ve = rss[k] / dfe;
                for (int b = 0; b < m; b++) {
                    tbg[(b * K + k) * 2 + 0] = beta[k][b];
                    tbg[(b * K + k) * 2 + 1] = ve * dg[b];
                }
            }
        }
    }
    return PG_LOCUS_OK;
}

// pearsons_correlation with missing frequencies (src/gwas/correlation_test.rs:21-31): pools whose frequency is NaN
// are dropped pairwise, the means and centred sums run over the remaining pools, n stays the full pool count.
template <int A, int K>
__device__ __noinline__ void explicit_corr_nan(const ScanParams &p, int64_t locus, unsigned kept, int m, const int *cols,
                                               const double *ys, double *tbg) {
    const Layout &lay = p.lay;
    const double *fl = p.freq + (size_t)locus * lay.freq_stride();
    const uint32_t *dl = p.depth + (size_t)locus * lay.depth_stride();
    double sx[PG_MAX_SLOTS], sy[K], sxx[PG_MAX_SLOTS], syy[K], sxy[K][PG_MAX_SLOTS];
    double cnt = 0.0;
    for (int pass = 0; pass < 2; pass++) {
        for (int a = 0; a < PG_MAX_SLOTS; a++) {
            if (pass == 0) sx[a] = 0.0; else sx[a] = sx[a] / cnt;
            sxx[a] = 0.0;
            for (int k = 0; k < K; k++) sxy[k][a] = 0.0;
        }
        for (int k = 0; k < K; k++) {
            if (pass == 0) sy[k] = 0.0; else sy[k] = sy[k] / cnt;
            syy[k] = 0.0;
        }
        int i0 = 0;
        for (int c = 0; c < lay.n_chunks; c++) {
            const int rcc = (c == lay.n_chunks - 1) ? lay.rc_last : lay.rc;
            const double *blk = fl + (size_t)c * lay.A * lay.rc;
            const int rows = min(rcc, lay.n - i0);
            for (int r = 0; r < rows; r++) {
                double f[A], F[A];
#pragma unroll
                for (int j = 0; j < A; j++) f[j] = blk[(size_t)j * rcc + r];
                renorm_row<A>(f, dl[i0 + r], kept, F);
                double x[PG_MAX_SLOTS];
                bool valid = true;
                for (int a = 0; a < m; a++) {
                    double v = 0.0;
#pragma unroll
                    for (int j = 0; j < A; j++)
                        if (j == cols[a]) v = F[j];
                    x[a] = v;
                    valid &= (v == v);
                }
                if (!valid) continue;
                if (pass == 0) {
                    cnt += 1.0;
                    for (int a = 0; a < m; a++) sx[a] += x[a];
                    for (int k = 0; k < K; k++) sy[k] += ys[k * lay.n_pad + i0 + r];
                } else {
                    for (int k = 0; k < K; k++) {
                        const double dy = ys[k * lay.n_pad + i0 + r] - sy[k];
                        syy[k] = fma(dy, dy, syy[k]);
                        for (int a = 0; a < m; a++) sxy[k][a] = fma(x[a] - sx[a], dy, sxy[k][a]);
                    }
                    for (int a = 0; a < m; a++) sxx[a] = fma(x[a] - sx[a], x[a] - sx[a], sxx[a]);
                }
            }
            i0 += rcc;
        }
    }
    for (int a = 0; a < m; a++)
        for (int k = 0; k < K; k++) {
            tbg[(a * K + k) * 2 + 0] = sxy[k][a] / (sqrt(sxx[a]) * sqrt(syy[k]));
            tbg[(a * K + k) * 2 + 1] = 0.0;
        }
}

// ---- end of a locus (all lanes, warp uniform): keep-mask and allele order from the reduced sums ---------------
template <int A, int K, bool W>
__device__ __noinline__ uint64_t decide_locus(const ScanParams &p, int64_t locus, double *tg, double *red, unsigned dm,
                                              const double *ys, const double *ws, int lane, double tol_rel) {
    using AC = Acc<A, K, W>;
    int status = PG_LOCUS_OK;
    unsigned kept = 0;
    bool slow = false;
    if ((double)dm < p.min_depth_f) {
        status = PG_LOCUS_FILTERED;  // sync.rs:217-229
    } else if (dm == 0u) {
        slow = true;  // a pool without coverage: NaN frequencies
    } else {
#pragma unroll
        for (int j = 0; j < A; j++) {
            double qj = W ? tg[AC::Q0 + j] : tg[AC::S0 + j] * p.w_uniform;
            const double tl = tol_rel * fmax(fabs(qj), 1.0);
            if (fabs(qj - p.maf) <= tl || fabs(qj - p.one_minus_maf) <= tl) qj = exact_q(p, locus, j, ws);
            if (!((qj < p.maf) | (qj > p.one_minus_maf))) kept |= 1u << j;
        }
        if (__popc(kept) < 2) {
            status = PG_LOCUS_FILTERED;  // sync.rs:284-286
        } else {
#pragma unroll
            for (int j = 0; j < A; j++)
                if (!((kept >> j) & 1u) && tg[AC::S0 + j] > 0.0) slow = true;  // a removed allele carries reads
        }
    }
    if (slow) status = slow_locus<A, K, W>(p, locus, ys, ws, red, tg, lane, dm == 0u, kept);
    int cols[PG_MAX_SLOTS] = {0, 0, 0, 0, 0};
    int nslots = 0;
    if (status == PG_LOCUS_OK) {
        nslots = __popc(kept) - 1;
        if (p.kind == PG_KIND_OLS) {
            // stable sort by decreasing column sum, drop the first (major) allele (sync.rs:478-505, ols.rs:227-230)
            double cs[A];
            bool tie = false;
#pragma unroll
            for (int j = 0; j < A; j++) cs[j] = tg[AC::S0 + j];
#pragma unroll
            for (int j = 0; j < A; j++)
#pragma unroll
                for (int l = j + 1; l < A; l++)
                    if (((kept >> j) & 1u) && ((kept >> l) & 1u) &&
                        fabs(cs[j] - cs[l]) <= tol_rel * fmax(fabs(cs[j]), fabs(cs[l])))
                        tie = true;
            if (tie) {
                double mine = 0.0;
                if (lane < A && ((kept >> lane) & 1u)) mine = exact_colsum<A>(p, locus, lane, kept);
#pragma unroll
                for (int j = 0; j < A; j++) cs[j] = __shfl_sync(PG_FULL_MASK, mine, j);
            }
#pragma unroll
            for (int j = 0; j < A; j++) {
                if (!((kept >> j) & 1u)) continue;
                int rank = 0;
#pragma unroll
                for (int l = 0; l < A; l++) {
                    if (l == j || !((kept >> l) & 1u)) continue;
                    if (cs[l] > cs[j] || (cs[l] == cs[j] && l < j)) rank++;
                }
#pragma unroll
                for (int s = 0; s < PG_MAX_SLOTS; s++)
                    if (rank == s + 1) cols[s] = j;
            }
        } else {
            // kept columns in file order, the last one is dropped (correlation_test.rs:94-98)
            int sidx = 0;
#pragma unroll
            for (int j = 0; j < A; j++) {
                if (!((kept >> j) & 1u)) continue;
#pragma unroll
                for (int ss = 0; ss < PG_MAX_SLOTS; ss++)
                    if (ss == sidx && sidx < nslots) cols[ss] = j;
                sidx++;
            }
        }
    }
    return pack_sel(status, nslots, cols, kept);
}

// Single-pass OLS from the reduced sums for M regressors, fully unrolled so that every matrix lives in registers
// (the generic chol_inv / solve_rhs above index local arrays dynamically and are kept for the two-pass fallback).
// Returns false when the centred X'X is not positive definite; sets redo when the single-pass form loses digits.
template <int M, int A, int K, bool W>
__device__ __forceinline__ bool ols_gram_m(const ScanParams &p, const double *tg, const int *cols, double nn,
                                           double *tbg, bool &redo) {
    using AC = Acc<A, K, W>;
    double sx[M], S[M][M], Lm[M][M], Li[M][M], dg[M], rinv[M];
    int c[M];
#pragma unroll
    for (int a = 0; a < M; a++) {
        c[a] = cols[a];
        sx[a] = tg[AC::S0 + c[a]];
    }
    const double inv_n = 1.0 / nn;
    double amp = 1.0;
#pragma unroll
    for (int a = 0; a < M; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) {
            const int lo = c[a] < c[b] ? c[a] : c[b], hi = c[a] < c[b] ? c[b] : c[a];
            const double raw = tg[AC::tri(lo, hi)];
            S[a][b] = raw - sx[a] * sx[b] * inv_n;
            if (a == b) amp = fmax(amp, raw / S[a][a]);
        }
    bool ok = true;
#pragma unroll
    for (int a = 0; a < M; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) {
            double v = S[a][b];
#pragma unroll
            for (int q = 0; q < b; q++) v -= Lm[a][q] * Lm[b][q];
            if (a == b) {
                if (!(v > 0.0)) {
                    ok = false;
                    v = 1.0;
                }
                Lm[a][a] = sqrt(v);
                rinv[a] = 1.0 / Lm[a][a];
            } else {
                Lm[a][b] = v * rinv[b];
            }
        }
#pragma unroll
    for (int a = 0; a < M; a++) {
        Li[a][a] = rinv[a];
#pragma unroll
        for (int b = 0; b < a; b++) {
            double v = 0.0;
#pragma unroll
            for (int q = b; q < a; q++) v -= Lm[a][q] * Li[q][b];
            Li[a][b] = v * rinv[a];
        }
    }
    double vif = 1.0;
#pragma unroll
    for (int b = 0; b < M; b++) {
        double v = 0.0;
#pragma unroll
        for (int a = b; a < M; a++) v += Li[a][b] * Li[a][b];
        dg[b] = v;
        vif = fmax(vif, S[b][b] * v);
    }
    redo = !ok || !(amp > 0.0);
    if (!ok) return false;
    const double inv_dfe = 1.0 / (nn - (double)(M + 1));
    const bool saturated = !(nn - (double)(M + 1) > 0.0);
#pragma unroll
    for (int k = 0; k < K; k++) {
        double z[M], zz = 0.0;
        const double ys_n = p.ysum[k] * inv_n;
#pragma unroll
        for (int a = 0; a < M; a++) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b <= a; b++) v += Li[a][b] * (tg[AC::C0 + c[b] * K + k] - sx[b] * ys_n);
            z[a] = v;
            zz += v * v;
        }
        double rss = p.syy[k] - zz;
        // digits lost by the single-pass form: centring (amp), collinearity (vif) and syy - zz
        if (!(64.0 * kEps * (amp * vif * zz + p.syy[k]) <= 1e-10 * rss)) redo = true;
        if (rss < 0.0) rss = 0.0;
        const double ve = saturated ? rss / (nn - (double)(M + 1)) : rss * inv_dfe;
#pragma unroll
        for (int b = 0; b < M; b++) {
            double v = 0.0;
#pragma unroll
            for (int a = b; a < M; a++) v += Li[a][b] * z[a];
            tbg[(b * K + k) * 2 + 0] = v;
            tbg[(b * K + k) * 2 + 1] = ve * dg[b];
        }
    }
    return true;
}

// ---- phase 2 (one lane per locus): centred normal equations -> (beta | r, var) per (allele slot, phenotype) -------
template <int A, int K, bool W>
__device__ __noinline__ uint64_t solve_locus(const ScanParams &p, int64_t locus, const double *tg, uint64_t sv,
                                             const double *ys, double *tbg) {
    using AC = Acc<A, K, W>;
    const Layout &lay = p.lay;
    const double nn = (double)lay.n;
    int status = (int)(sv & 0xff);
    const int m = (int)((sv >> 8) & 0xff);
    int cols[PG_MAX_SLOTS];
#pragma unroll
    for (int s = 0; s < PG_MAX_SLOTS; s++) cols[s] = (int)((sv >> (16 + 4 * s)) & 0xf);
    double fmean[PG_MAX_SLOTS];
    if (status == PG_LOCUS_OK) {
        double sx[PG_MAX_SLOTS], xbar[PG_MAX_SLOTS];
        bool has_nan = false;
        for (int a = 0; a < m; a++) {
            sx[a] = tg[AC::S0 + cols[a]];
            const double pjj = tg[AC::tri(cols[a], cols[a])];
            has_nan |= (pjj != pjj);
            xbar[a] = sx[a] / nn;
            fmean[a] = (pjj != pjj) ? nan("") : xbar[a];
        }
        const unsigned keptm = (unsigned)((sv >> 40) & 0x3f);
        if (p.kind == PG_KIND_OLS) {
            if (lay.n < m + 1) {
                status = PG_LOCUS_UNSUPPORTED;
            } else if (has_nan) {
                for (int i = 0; i < m * K * 2; i++) tbg[i] = nan("");
            } else {
                bool redo = false;
                switch (m) {
                    case 1: ols_gram_m<1, A, K, W>(p, tg, cols, nn, tbg, redo); break;
                    case 2:
                        if constexpr (A >= 3) ols_gram_m<2, A, K, W>(p, tg, cols, nn, tbg, redo);
                        break;
                    case 3:
                        if constexpr (A >= 4) ols_gram_m<3, A, K, W>(p, tg, cols, nn, tbg, redo);
                        break;
                    case 4:
                        if constexpr (A >= 5) ols_gram_m<4, A, K, W>(p, tg, cols, nn, tbg, redo);
                        break;
                    default:
                        if constexpr (A >= 6) ols_gram_m<5, A, K, W>(p, tg, cols, nn, tbg, redo);
                        break;
                }
                if (redo) status = explicit_locus<A, K>(p, locus, keptm, m, cols, xbar, ys, tbg);
            }
        } else {
            bool redo = false;
            for (int a = 0; a < m; a++) {
                const double raw = tg[AC::tri(cols[a], cols[a])];
                const double sxx = raw - sx[a] * sx[a] / nn;
                if (!(raw <= 1e4 * sxx)) redo = true;
                for (int k = 0; k < K; k++) {
                    const double sxy = tg[AC::C0 + cols[a] * K + k] - sx[a] * p.ysum[k] / nn;
                    tbg[(a * K + k) * 2 + 0] = sxy / (sqrt(sxx) * sqrt(p.syy[k]));
                    tbg[(a * K + k) * 2 + 1] = 0.0;
                }
            }
            if (has_nan)
                explicit_corr_nan<A, K>(p, locus, keptm, m, cols, ys, tbg);
            else if (redo)
                explicit_locus<A, K>(p, locus, keptm, m, cols, xbar, ys, tbg);
        }
        if (status != PG_LOCUS_OK) sv = (sv & ~(uint64_t)0xff) | (uint64_t)status;
    }
    if (p.write_meta) {
        uint64_t mv = (uint64_t)status;
        if (status == PG_LOCUS_OK) {
            mv |= (uint64_t)m << 8;
            for (int s = 0; s < m; s++) mv |= (uint64_t)p.codes[cols[s]] << (16 + 8 * s);
        }
        p.meta[locus] = mv;
        for (int s = 0; s < A - 1; s++)
            p.freq_mean[(size_t)locus * (A - 1) + s] = (status == PG_LOCUS_OK && s < m) ? fmean[s] : nan("");
    }
    return sv;
}

// ---- phase 3 (one lane per (locus, allele slot, phenotype)): t and p --------------------------------------------
static __device__ __noinline__ void finish_task(const ScanParams &p, const PTableDev &ptab, bool valid, double v0, double v1,
                                         double *o) {
    const double nn = (double)p.lay.n;
    double o0 = nan(""), o1 = nan(""), o2 = nan(""), o3 = nan("");
    if (valid) {
        if (p.kind == PG_KIND_OLS) {
            // estimate_significance, src/gwas/ols.rs:139-154
            const double se = sqrt(v1);
            const double tt = (fabs(v0) <= kEps) ? 0.0 : v0 / se;
            double pv;
            if (fabs(tt) <= kEps || tt != tt)
                pv = 1.0;
            else
                pv = p.ptab ? student_two_sided_tab(fabs(tt), p.df, ptab) : student_two_sided(fabs(tt), p.df, p.ln_beta);
            o0 = v0;
            o1 = se;
            o2 = tt;
            o3 = pv;
        } else {
            // pearsons_correlation, src/gwas/correlation_test.rs:52-70
            const double r = v0;
            if (r == r) {
                const double s2 = (1.0 - r * r) / (nn - 2.0);
                o1 = r;
                if (s2 <= 0.0) {
                    o0 = r;
                    o3 = kEps;
                } else {
                    const double tt = r / sqrt(s2);
                    o2 = tt;
                    o3 = (p.lay.n > 2) ? (p.ptab ? student_two_sided_tab(fabs(tt), p.df, ptab)
                                                 : student_two_sided(fabs(tt), p.df, p.ln_beta))
                                       : nan("");
                    o0 = round(r * 1e7) / 1e7;
                }
            }
        }
    }
    *reinterpret_cast<double2 *>(o) = make_double2(o0, o1);
    *reinterpret_cast<double2 *>(o + 2) = make_double2(o2, o3);
}

constexpr int kMaxWarps = 16;

template <int A, int K, bool W>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) scan_kernel(const ScanParams p) {
    using AC = Acc<A, K, W>;
    using WS = WarpSmem<A, K, W>;
    constexpr int T = WS::T;
    constexpr int G = WS::G;
    extern __shared__ __align__(128) unsigned char smem[];
    const Layout lay = p.lay;
    const int n_pad = lay.n_pad;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double *ys = reinterpret_cast<double *>(smem);
    double *ws = W ? ys + (size_t)K * n_pad : nullptr;
    unsigned char *wb = smem + scan_common_bytes(K, n_pad, W) + (size_t)warp * WS::bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wb);
    double *fbuf = reinterpret_cast<double *>(wb + WS::off_f);
    uint32_t *dbuf = reinterpret_cast<uint32_t *>(wb + WS::off_d);
    double *tot = reinterpret_cast<double *>(wb + WS::off_tot);
    uint64_t *sel = reinterpret_cast<uint64_t *>(wb + WS::off_sel);
    double *tb = reinterpret_cast<double *>(wb + WS::off_tb);

    for (int i = threadIdx.x; i < K * n_pad; i += blockDim.x) ys[i] = p.yc[i];
    if (W)
        for (int i = threadIdx.x; i < n_pad; i += blockDim.x) ws[i] = p.w[i];
    if (lane == 0) {
        for (int b = 0; b < kNBuf; b++) mbar_init(&bars[b], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t L = p.n_loci;
    const int64_t NG = (L + G - 1) / G;
    const int64_t gw = (int64_t)blockIdx.x * nwarps + warp;  // consecutive warps take consecutive groups
    const int64_t TW = (int64_t)gridDim.x * nwarps;
    if (gw >= NG) return;
    const int n_chunks = lay.n_chunks;
    const size_t fstride = lay.freq_stride(), dstride = lay.depth_stride();
    const double tol_rel = 2.0 * ((double)lay.n + 8.0) * kEps;
    const PTableDev ptab = {reinterpret_cast<const double4 *>(p.ptab), p.ptab_vmax, p.ptab_inv_h, p.ptab_M};

    // producer side of the warp's private ring: the tile stream is (group, locus in group, chunk), and only the
    // last group of the job can be short, so the stream of a warp simply ends at the first locus >= L.  All of the
    // iterator state is plain scalars so that it stays in registers.
    int64_t i_group = gw, i_locus = gw * G;
    int i_li = 0, i_chunk = 0;
    bool i_done = false;
#define PG_ISSUE_NEXT(BUF)                                                                                  \
    do {                                                                                                    \
        if (!i_done) {                                                                                      \
            if (lane == 0) {                                                                                \
                const int rcc_ = (i_chunk == n_chunks - 1) ? lay.rc_last : lay.rc;                          \
                const uint32_t fbytes_ = (uint32_t)(A * rcc_ * 8), dbytes_ = (uint32_t)(rcc_ * 4);          \
                const double *fsrc_ = p.freq + (size_t)i_locus * fstride + (size_t)i_chunk * A * lay.rc;    \
                const uint32_t *dsrc_ = p.depth + (size_t)i_locus * dstride + (size_t)i_chunk * lay.rc;     \
                fence_proxy_async(); /* the buffer may have served as reduction scratch (generic writes) */ \
                mbar_expect_tx(&bars[BUF], fbytes_ + dbytes_);                                              \
                bulk_g2s(fbuf + (size_t)(BUF) * (WS::fbuf_bytes / 8), fsrc_, fbytes_, &bars[BUF]);          \
                bulk_g2s(dbuf + (size_t)(BUF) * (WS::dbuf_bytes / 4), dsrc_, dbytes_, &bars[BUF]);          \
            }                                                                                               \
            if (++i_chunk == n_chunks) {                                                                    \
                i_chunk = 0;                                                                                \
                i_locus++;                                                                                  \
                if (++i_li == G) {                                                                          \
                    i_li = 0;                                                                               \
                    i_group += TW;                                                                          \
                    i_locus = i_group * G;                                                                  \
                    if (i_group >= NG) i_done = true;                                                       \
                }                                                                                           \
                if (i_locus >= L) i_done = true;                                                            \
            }                                                                                               \
        }                                                                                                   \
    } while (0)
#pragma unroll
    for (int b = 0; b < kNBuf; b++) PG_ISSUE_NEXT(b);

    uint32_t phase_bits = 0;  // bit b = parity to wait for on buffer b
    int buf = 0;
    for (int64_t group = gw; group < NG; group += TW) {
        const int64_t gl0 = group * G;
        const int gn = (int)min((int64_t)G, L - gl0);  // loci of this group
        for (int g = 0; g < gn; g++) {
            const int64_t locus = gl0 + g;
            double acc[AC::NP];
#pragma unroll
            for (int i = 0; i < AC::NP; i++) acc[i] = 0.0;
            unsigned dmin = 0xFFFFFFFFu;
            double *red = nullptr;
            for (int chunk = 0; chunk < n_chunks; chunk++) {
                const int rcc = (chunk == n_chunks - 1) ? lay.rc_last : lay.rc;
                const int row0 = chunk * lay.rc;
                mbar_wait(&bars[buf], (phase_bits >> buf) & 1u);
                phase_bits ^= (1u << buf);
                const double *fb = fbuf + (size_t)buf * (WS::fbuf_bytes / 8);
                const uint32_t *db = dbuf + (size_t)buf * (WS::dbuf_bytes / 4);
                // ---- phase 1: lane handles rows 2*lane, 2*lane+1 (+64 ...) with 128-bit shared loads
                for (int r = 2 * lane; r < ((p.debug & 1) ? 0 : rcc); r += 64) {
                    double2 f2[A];
#pragma unroll
                    for (int j = 0; j < A; j++) f2[j] = *reinterpret_cast<const double2 *>(fb + (size_t)j * rcc + r);
                    const uint2 d2 = *reinterpret_cast<const uint2 *>(db + r);
                    double2 y2[K];
#pragma unroll
                    for (int k = 0; k < K; k++)
                        y2[k] = *reinterpret_cast<const double2 *>(ys + (size_t)k * n_pad + row0 + r);
                    double2 w2 = make_double2(0.0, 0.0);
                    if (W) w2 = *reinterpret_cast<const double2 *>(ws + row0 + r);
                    dmin = min(dmin, min(d2.x, d2.y));
                    double fa[A], ya[K];
#pragma unroll
                    for (int j = 0; j < A; j++) fa[j] = f2[j].x;
#pragma unroll
                    for (int k = 0; k < K; k++) ya[k] = y2[k].x;
                    accum_row<A, K, W, false>(acc, fa, ya, w2.x);
#pragma unroll
                    for (int j = 0; j < A; j++) fa[j] = f2[j].y;
#pragma unroll
                    for (int k = 0; k < K; k++) ya[k] = y2[k].y;
                    accum_row<A, K, W, false>(acc, fa, ya, w2.y);
                }
                __syncwarp();
                if (chunk == n_chunks - 1 && !(p.debug & 2)) {
                    // the buffer just consumed doubles as the reduction scratch before it is refilled
                    red = const_cast<double *>(fb);
                    double *tg = tot + (size_t)g * AC::NP;
                    const unsigned dm = __reduce_min_sync(PG_FULL_MASK, dmin);
                    reduce_store<AC::N, AC::NP, WS::RED_ROWS>(acc, red, tg, lane);
                    __syncwarp();
                    const uint64_t sv = decide_locus<A, K, W>(p, locus, tg, red, dm, ys, ws, lane, tol_rel);
                    if (lane == 0) sel[g] = sv;
                    __syncwarp();
                }
                PG_ISSUE_NEXT(buf);  // refill this buffer with the tile kNBuf ahead
                buf = (buf + 1 == kNBuf) ? 0 : buf + 1;
            }
        }
        // ---- phase 2: lane = locus of the group
        __syncwarp();
        if (p.debug & 4) continue;
        if (lane < gn) {
            const uint64_t sv = solve_locus<A, K, W>(p, gl0 + lane, tot + (size_t)lane * AC::NP, sel[lane], ys,
                                                     tb + (size_t)lane * T * 2);
            sel[lane] = sv;
        }
        __syncwarp();
        // ---- phase 3: lane = (locus, allele slot, phenotype)
        for (int task = lane; task < ((p.debug & 8) ? 0 : gn * T); task += 32) {
            const int g = task / T, rem = task - g * T;
            const int slot = rem / K, kk = rem - slot * K;
            const uint64_t sv = sel[g];
            const bool valid = ((int)(sv & 0xff) == PG_LOCUS_OK) && slot < (int)((sv >> 8) & 0xff);
            const double v0 = tb[((size_t)g * T + slot * K + kk) * 2 + 0];
            const double v1 = tb[((size_t)g * T + slot * K + kk) * 2 + 1];
            double *o = p.stats + (((size_t)(gl0 + g) * (A - 1) + slot) * p.k_total + p.phen_base + kk) * 4;
            finish_task(p, ptab, valid, v0, v1, o);
        }
        __syncwarp();
    }
#undef PG_ISSUE_NEXT
}

template <int A, int K, bool W>
cudaError_t launch_scan_t(const ScanParams &p, int sm_count, cudaStream_t s) {
    using WS = WarpSmem<A, K, W>;
    const size_t common = scan_common_bytes(K, p.lay.n_pad, W);
    const size_t avail = 227 * 1024;
    if (common + WS::bytes > avail) return cudaErrorInvalidConfiguration;
    int nwarps = (int)((avail - common) / WS::bytes);
    if (nwarps > kMaxWarps) nwarps = kMaxWarps;
    const size_t smem = common + (size_t)nwarps * WS::bytes;
    auto kern = scan_kernel<A, K, W>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t NG = (p.n_loci + WS::G - 1) / WS::G;
    int64_t ctas = (NG + nwarps - 1) / nwarps;
    if (ctas > sm_count) ctas = sm_count;
    if (ctas < 1) ctas = 1;
    kern<<<(unsigned)ctas, nwarps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

template <int A>
cudaError_t launch_scan_a(const ScanParams &p, int sm_count, cudaStream_t s) {
    const bool w = p.weighted != 0;
    switch (p.K) {
        case 1: return w ? launch_scan_t<A, 1, true>(p, sm_count, s) : launch_scan_t<A, 1, false>(p, sm_count, s);
        case 2: return w ? launch_scan_t<A, 2, true>(p, sm_count, s) : launch_scan_t<A, 2, false>(p, sm_count, s);
        case 3: return w ? launch_scan_t<A, 3, true>(p, sm_count, s) : launch_scan_t<A, 3, false>(p, sm_count, s);
        case 4: return w ? launch_scan_t<A, 4, true>(p, sm_count, s) : launch_scan_t<A, 4, false>(p, sm_count, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace pg
