// pg_ingest.cu -- ingestion kernels: parsed counts (or a host-built frequency matrix) -> the chunked
// f64 first-stage frequency matrix + depth vector the scan kernel streams, and the synthetic workload
// generator.  Frequencies are one IEEE division of exactly representable integers, i.e. bit-identical
// to LocusCounts::to_frequencies (src/base/sync.rs:166-192).  Every ingest also leaves dmin[locus] = the smallest
// pool depth of the locus and the renormalisation hint of the scan.
#include <stdlib.h>

#include "pg_device.cuh"
#include "pg_internal.h"

namespace pg {

// hint of a locus: bit 7 = "renormalise over the alleles in bits 0..5": every pool has coverage, the depth filter
// passes, at least two alleles survive the MAF filter, a removed allele carries reads (the regression then runs on
// c_ij / sum_kept c_i., src/base/sync.rs:166-192 after the filter) and every pooled frequency is further from both
// thresholds than twice the rounding bound of ANY summation order -- so the keep-mask in the hint is the reference's.
// The streaming scan renormalises such loci on the fly; every other locus gets 0 and takes the scan's own decision
// path (which defers what it cannot settle to the fix-up kernel).
struct HintParams {
    double maf, one_minus_maf, min_depth_f;
    int enabled;
};
// alleles whose pooled frequency sits within twice the rounding bound of any summation order of a MAF threshold
__device__ __forceinline__ unsigned ambiguous_alleles(const double *q, int A, int n, const HintParams &hp) {
    const double tol_rel = 4.0 * ((double)n + 8.0) * kEps;
    unsigned amb = 0;
    for (int j = 0; j < A; j++) {
        const double tl = tol_rel * fmax(fabs(q[j]), 1.0);
        if (fabs(q[j] - hp.maf) <= tl || fabs(q[j] - hp.one_minus_maf) <= tl) amb |= 1u << j;
    }
    return amb;
}

// `exact` = alleles whose q was re-evaluated in the reference's own order (any distance from a threshold is decisive)
__device__ __forceinline__ uint8_t locus_hint(const double *q, int A, int n, uint32_t dm, const HintParams &hp,
                                              unsigned exact) {
    if (!hp.enabled) return 0;
    const double tol_rel = 4.0 * ((double)n + 8.0) * kEps;
    unsigned kept = 0;
    bool removed_with_reads = false, safe = true;
    for (int j = 0; j < A; j++) {
        const double tl = tol_rel * fmax(fabs(q[j]), 1.0);
        if (!((exact >> j) & 1u) && (fabs(q[j] - hp.maf) <= tl || fabs(q[j] - hp.one_minus_maf) <= tl)) safe = false;
        if (!((q[j] < hp.maf) | (q[j] > hp.one_minus_maf)))
            kept |= 1u << j;
        else if (q[j] > 0.0)
            removed_with_reads = true;
    }
    const bool depth_ok = dm != 0u && !((double)dm < hp.min_depth_f);
    return (safe && depth_ok && removed_with_reads && __popc(kept) >= 2) ? (uint8_t)(0x80u | kept) : (uint8_t)0;
}

// A group of P lanes (P = 32 for more than 16 pools, else the next power of two) owns a locus: every lane converts
// the pools q0, q0 + P, ... and keeps the pooled frequencies q_j = sum_i f_ij w_i (NaN skipped) and the smallest
// depth of its pools; a butterfly over the group completes them, its first lane stores dmin and the hint.  No
// atomics, no scratch, and a locus gets the same bits wherever it lands.
template <typename RowFn>
__device__ __forceinline__ void ingest_loci(int64_t n_loci, const Layout &lay, int P, const double *freq,
                                            const double *__restrict__ w, uint32_t *__restrict__ dmin,
                                            uint8_t *__restrict__ hint, const HintParams &hp, RowFn &&row) {
    const int lane = threadIdx.x & 31, sub = lane / P, q0 = lane % P, lpw = 32 / P;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t l0 = gw * lpw; l0 < n_loci; l0 += nw * lpw) {
        const int64_t locus = l0 + sub;
        double q[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
        uint32_t dm = 0xFFFFFFFFu;
        if (locus < n_loci)
            for (int i = q0; i < lay.n_pad; i += P) row(locus, i, q, dm);
        for (int off = P >> 1; off >= 1; off >>= 1) {
            for (int j = 0; j < lay.A; j++) q[j] += __shfl_xor_sync(0xFFFFFFFFu, q[j], off);
            dm = min(dm, __shfl_xor_sync(0xFFFFFFFFu, dm, off));
        }
        // a pooled frequency within rounding of a threshold (one read in one pool of depth 10 among 100 pools IS
        // 0.001): the group re-evaluates it in the reference's order -- pool by pool, separately rounded multiply and
        // add, src/base/sync.rs:258-271 -- from the frequencies it has just written, so the hint can still vouch for
        // the keep-mask and the locus stays out of the fix-up kernel
        unsigned amb = (hp.enabled && locus < n_loci && dm != 0u) ? ambiguous_alleles(q, lay.A, lay.n, hp) : 0u;
        if (__any_sync(0xFFFFFFFFu, amb != 0u)) {
            __syncwarp();  // the group's stores to freq are visible to its lanes
            const double *fl = freq + (size_t)(locus < n_loci ? locus : 0) * lay.freq_stride();
            for (int j = 0; j < lay.A; j++) {
                if (!__any_sync(0xFFFFFFFFu, (amb >> j) & 1u)) continue;
                double e = 0.0;
                for (int r0 = 0; r0 < lay.n; r0 += P) {
                    const int i = r0 + q0;
                    double term = 0.0;
                    if (i < lay.n && locus < n_loci) {
                        const double f = __ldcg(fl + lay.freq_off(i, j));  // written by another lane of this warp
                        term = (f != f) ? 0.0 : __dmul_rn(f, w[i]);
                    }
                    const int m = min(P, lay.n - r0);
                    for (int t = 0; t < m; t++) e = __dadd_rn(e, __shfl_sync(0xFFFFFFFFu, term, t, P));
                }
                if ((amb >> j) & 1u) q[j] = e;
            }
        }
        if (locus < n_loci && q0 == 0) {
            dmin[locus] = dm;
            hint[locus] = locus_hint(q, lay.A, lay.n, dm, hp, amb);
        }
    }
}

template <typename CT>
__global__ void __launch_bounds__(256) ingest_counts_kernel(const CT *__restrict__ counts, int64_t n_loci, int n,
                                                            int A_in, int drop_col, Layout lay, int P,
                                                            double *__restrict__ freq, uint32_t *__restrict__ depth,
                                                            uint32_t *__restrict__ dmin, uint8_t *__restrict__ hint,
                                                            const double *__restrict__ w, HintParams hp) {
    ingest_loci(n_loci, lay, P, freq, w, dmin, hint, hp, [&](int64_t locus, int i, double *q, uint32_t &dm) {
        double *fl = freq + (size_t)locus * lay.freq_stride();
        if (i >= n) {  // padding row
            for (int j = 0; j < lay.A; j++) fl[lay.freq_off(i, j)] = 0.0;
            depth[(size_t)locus * lay.n_pad + i] = 0xFFFFFFFFu;
            return;
        }
        const CT *cl = counts + (size_t)locus * A_in * n + i;
        uint32_t c[PG_MAX_ALLELES];
        uint64_t d = 0;
        int jj = 0;
        for (int j = 0; j < A_in; j++) {
            if (j == drop_col) continue;
            c[jj] = (uint32_t)cl[(size_t)j * n];
            d += c[jj];
            jj++;
        }
        if (d > 0xFFFFFFFEull) d = 0xFFFFFFFEull;  // documented limit: per-pool depth < 2^32 - 1
        // f = c / d, correctly rounded like the reference's division, at one reciprocal per pool instead of one division
        // per allele: with r = RN(1 / d), q = RN(c r) and the exact remainder c - q d, RN(q + rem r) IS RN(c / d)
        // (Markstein's correction step; bit-identical for all 32-bit operands -- a host-side test under tests/ checks
        // the same sequence on the host exhaustively for d <= 4096 and on random pairs)
        const double dd = (double)d, wi = w[i];
        const double rd = __drcp_rn(dd);
        for (int j = 0; j < lay.A; j++) {
            const double cj = (double)c[j];
            const double q0 = __dmul_rn(cj, rd);
            const double f = (d == 0) ? nan("") : __fma_rn(__fma_rn(-q0, dd, cj), rd, q0);
            fl[lay.freq_off(i, j)] = f;
            if (d != 0) q[j] = fma(f, wi, q[j]);
        }
        depth[(size_t)locus * lay.n_pad + i] = (uint32_t)d;
        dm = min(dm, (uint32_t)d);
    });
}

__global__ void __launch_bounds__(256) ingest_freq_kernel(const double *__restrict__ fin, const uint32_t *__restrict__ din,
                                                          int64_t n_loci, int n, Layout lay, int P,
                                                          double *__restrict__ freq, uint32_t *__restrict__ depth,
                                                          uint32_t *__restrict__ dmin, uint8_t *__restrict__ hint,
                                                          const double *__restrict__ w, HintParams hp) {
    ingest_loci(n_loci, lay, P, freq, w, dmin, hint, hp, [&](int64_t locus, int i, double *q, uint32_t &dm) {
        double *fl = freq + (size_t)locus * lay.freq_stride();
        const bool pad = i >= n;
        for (int j = 0; j < lay.A; j++) {
            const double f = pad ? 0.0 : fin[((size_t)locus * lay.A + j) * n + i];
            fl[lay.freq_off(i, j)] = f;
            if (!pad && f == f) q[j] = fma(f, w[i], q[j]);
        }
        const uint32_t d = pad ? 0xFFFFFFFFu : din[(size_t)locus * n + i];
        depth[(size_t)locus * lay.n_pad + i] = d;
        dm = min(dm, d);
    });
}

__global__ void __launch_bounds__(256) synth_kernel(uint64_t seed, int64_t first_locus, int64_t n_loci, int n,
                                                    int A_in, uint32_t *__restrict__ counts) {
    const int64_t total = n_loci * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = idx / n;
        const int i = (int)(idx - l * n);
        uint32_t c[PG_MAX_ALLELES];
        synth_counts(seed, first_locus + l, i, A_in, c);
        for (int a = 0; a < A_in; a++) counts[((size_t)l * A_in + a) * n + i] = c[a];
    }
}

static int grid_for(int64_t total) {
    int64_t b = (total + 255) / 256;
    const int64_t cap = 148 * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// lanes per locus of the ingest kernels
static int ingest_group(const Layout &lay) {
    int P = 32;
    while (P > 1 && (P >> 1) >= lay.n_pad) P >>= 1;
    return P;
}
static HintParams hint_params(const IngestOut &o) {
    static const bool no_hint = getenv("PG_NOHINT") != nullptr;  // tuning knob: every locus takes the scan's own path
    HintParams hp;
    hp.maf = o.maf;
    hp.one_minus_maf = o.one_minus_maf;
    hp.min_depth_f = o.min_depth_f;
    hp.enabled = no_hint ? 0 : 1;
    return hp;
}
template <typename CT>
static cudaError_t launch_ingest_t(const CT *counts, int64_t n_loci, int n, int A_in, int drop_col, const Layout &lay,
                                   const IngestOut &o, cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    const int P = ingest_group(lay);
    ingest_counts_kernel<CT><<<grid_for(n_loci * P), 256, 0, s>>>(counts, n_loci, n, A_in, drop_col, lay, P, o.freq,
                                                                 o.depth, o.dmin, o.hint, o.w, hint_params(o));
    return cudaGetLastError();
}
cudaError_t launch_ingest_u32(const uint32_t *counts, int64_t n_loci, int n, int A_in, int drop_col, const Layout &lay,
                              const IngestOut &o, cudaStream_t s) {
    return launch_ingest_t<uint32_t>(counts, n_loci, n, A_in, drop_col, lay, o, s);
}
cudaError_t launch_ingest_u16(const uint16_t *counts, int64_t n_loci, int n, int A_in, int drop_col, const Layout &lay,
                              const IngestOut &o, cudaStream_t s) {
    return launch_ingest_t<uint16_t>(counts, n_loci, n, A_in, drop_col, lay, o, s);
}
cudaError_t launch_ingest_u8(const uint8_t *counts, int64_t n_loci, int n, int A_in, int drop_col, const Layout &lay,
                             const IngestOut &o, cudaStream_t s) {
    return launch_ingest_t<uint8_t>(counts, n_loci, n, A_in, drop_col, lay, o, s);
}
cudaError_t launch_ingest_freq(const double *fin, const uint32_t *din, int64_t n_loci, int n, const Layout &lay,
                               const IngestOut &o, cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    const int P = ingest_group(lay);
    ingest_freq_kernel<<<grid_for(n_loci * P), 256, 0, s>>>(fin, din, n_loci, n, lay, P, o.freq, o.depth, o.dmin, o.hint,
                                                            o.w, hint_params(o));
    return cudaGetLastError();
}
template <typename CT>
__global__ void __launch_bounds__(256) widen_kernel(const CT *__restrict__ src, uint32_t *__restrict__ dst, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (uint32_t)src[i];
}
cudaError_t launch_widen(const void *src, int elem_bytes, uint32_t *dst, size_t count, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    if (elem_bytes == 1)
        widen_kernel<uint8_t><<<grid_for((int64_t)count), 256, 0, s>>>((const uint8_t *)src, dst, count);
    else if (elem_bytes == 2)
        widen_kernel<uint16_t><<<grid_for((int64_t)count), 256, 0, s>>>((const uint16_t *)src, dst, count);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_synth(uint64_t seed, int64_t first_locus, int64_t n_loci, int n, int A_in, uint32_t *counts,
                         cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    synth_kernel<<<grid_for(n_loci * n), 256, 0, s>>>(seed, first_locus, n_loci, n, A_in, counts);
    return cudaGetLastError();
}

}  // namespace pg
