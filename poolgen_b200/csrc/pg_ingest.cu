// pg_ingest.cu -- ingestion kernels: parsed counts (or a host-built frequency matrix) -> the chunked
// f64 first-stage frequency matrix + depth vector the scan kernel streams, and the synthetic workload
// generator.  Frequencies are one IEEE division of exactly representable integers, i.e. bit-identical
// to LocusCounts::to_frequencies (src/base/sync.rs:166-192).  Every ingest also leaves dmin[locus] = the smallest
// pool depth of the locus (the caller presets dmin to 0xFFFFFFFF).
#include <stdlib.h>

#include "pg_device.cuh"
#include "pg_internal.h"

namespace pg {

// per-locus pooled frequencies q_j = sum_i f_ij w_i (NaN skipped): the lanes of a warp that hold the same locus are
// contiguous, so a segmented shuffle reduction leaves the sum of every run in its first lane, which issues ONE atomic
// per (warp, locus, allele).  Must be called by all 32 lanes at a convergent point (locus = -1 for lanes without a
// row).  The sums only feed the renormalisation HINT of the scan, never a decision on their own.
__device__ __forceinline__ void fold_q(double *qbuf, int64_t locus, int A, const double *fw) {
    const int lane = threadIdx.x & 31;
    const int64_t up = __shfl_up_sync(0xFFFFFFFFu, locus, 1);
    const bool head = lane == 0 || up != locus;
    bool same[5];
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int64_t other = __shfl_down_sync(0xFFFFFFFFu, locus, 1 << s);
        same[s] = (lane + (1 << s) < 32) && other == locus;
    }
    for (int j = 0; j < A; j++) {
        double v = fw[j];
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const double t = __shfl_down_sync(0xFFFFFFFFu, v, 1 << s);
            if (same[s]) v += t;
        }
        if (head && locus >= 0 && v != 0.0) atomicAdd(qbuf + (size_t)locus * A + j, v);
    }
}

// the same for the smallest depth of the locus
__device__ __forceinline__ void fold_dmin(uint32_t *dmin, int64_t locus, uint32_t d) {
    const int lane = threadIdx.x & 31;
    const int64_t up = __shfl_up_sync(0xFFFFFFFFu, locus, 1);
    const bool head = lane == 0 || up != locus;
    uint32_t v = d;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int64_t other = __shfl_down_sync(0xFFFFFFFFu, locus, 1 << s);
        const uint32_t t = __shfl_down_sync(0xFFFFFFFFu, v, 1 << s);
        if ((lane + (1 << s) < 32) && other == locus) v = min(v, t);
    }
    if (head && locus >= 0) atomicMin(dmin + locus, v);
}

// hint of a locus: bit 7 = "renormalise over the alleles in bits 0..5": every pool has coverage, the depth filter
// passes, at least two alleles survive the MAF filter, a removed allele carries reads (the regression then runs on
// c_ij / sum_kept c_i., src/base/sync.rs:166-192 after the filter) and every pooled frequency is further from both
// thresholds than twice the rounding bound of ANY summation order -- so the keep-mask in the hint is the reference's.
// The streaming scan renormalises such loci on the fly; every other locus gets 0 and takes the scan's own decision
// path (which defers what it cannot settle to the fix-up kernel).
__global__ void __launch_bounds__(256) hint_kernel(const double *__restrict__ qbuf, const uint32_t *__restrict__ dmin,
                                                   int64_t n_loci, int A, int n, double maf, double one_minus_maf,
                                                   double min_depth_f, uint8_t *__restrict__ hint) {
    const double tol_rel = 4.0 * ((double)n + 8.0) * kEps;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n_loci; l += (int64_t)gridDim.x * blockDim.x) {
        unsigned kept = 0;
        bool removed_with_reads = false, safe = true;
        for (int j = 0; j < A; j++) {
            const double q = qbuf[(size_t)l * A + j];
            const double tl = tol_rel * fmax(fabs(q), 1.0);
            if (fabs(q - maf) <= tl || fabs(q - one_minus_maf) <= tl) safe = false;
            if (!((q < maf) | (q > one_minus_maf)))
                kept |= 1u << j;
            else if (q > 0.0)
                removed_with_reads = true;
        }
        const uint32_t dm = dmin[l];
        const bool depth_ok = dm != 0u && !((double)dm < min_depth_f);
        hint[l] = (safe && depth_ok && removed_with_reads && __popc(kept) >= 2) ? (uint8_t)(0x80u | kept) : (uint8_t)0;
    }
}

template <typename CT>
__global__ void __launch_bounds__(256) ingest_counts_kernel(const CT *__restrict__ counts, int64_t n_loci, int n,
                                                            int A_in, int drop_col, Layout lay,
                                                            double *__restrict__ freq, uint32_t *__restrict__ depth,
                                                            uint32_t *__restrict__ dmin, double *__restrict__ qbuf,
                                                            const double *__restrict__ w) {
    const int64_t total = n_loci * lay.n_pad;
    // the whole warp stays in the loop: the folds at the end shuffle across all 32 lanes
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx - (int64_t)(threadIdx.x & 31) < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t locus = -1;
        uint32_t dfold = 0xFFFFFFFFu;
        double fw[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
        if (idx < total) {
            locus = idx / lay.n_pad;
            const int i = (int)(idx - locus * lay.n_pad);
            double *fl = freq + (size_t)locus * lay.freq_stride();
            if (i >= n) {  // padding row
                for (int j = 0; j < lay.A; j++) fl[lay.freq_off(i, j)] = 0.0;
                depth[(size_t)locus * lay.n_pad + i] = 0xFFFFFFFFu;
            } else {
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
                const double dd = (double)d, wi = w[i];
                for (int j = 0; j < lay.A; j++) {
                    const double f = (d == 0) ? nan("") : (double)c[j] / dd;
                    fl[lay.freq_off(i, j)] = f;
                    fw[j] = (d == 0) ? 0.0 : f * wi;
                }
                depth[(size_t)locus * lay.n_pad + i] = (uint32_t)d;
                dfold = (uint32_t)d;
            }
        }
        fold_dmin(dmin, locus, dfold);
        fold_q(qbuf, locus, lay.A, fw);
    }
}

__global__ void __launch_bounds__(256) ingest_freq_kernel(const double *__restrict__ fin, const uint32_t *__restrict__ din,
                                                          int64_t n_loci, int n, Layout lay,
                                                          double *__restrict__ freq, uint32_t *__restrict__ depth,
                                                          uint32_t *__restrict__ dmin, double *__restrict__ qbuf,
                                                          const double *__restrict__ w) {
    const int64_t total = n_loci * lay.n_pad;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx - (int64_t)(threadIdx.x & 31) < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t locus = -1;
        uint32_t dfold = 0xFFFFFFFFu;
        double fw[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
        if (idx < total) {
            locus = idx / lay.n_pad;
            const int i = (int)(idx - locus * lay.n_pad);
            double *fl = freq + (size_t)locus * lay.freq_stride();
            const bool pad = i >= n;
            for (int j = 0; j < lay.A; j++) {
                const double f = pad ? 0.0 : fin[((size_t)locus * lay.A + j) * n + i];
                fl[lay.freq_off(i, j)] = f;
                fw[j] = (pad || f != f) ? 0.0 : f * w[i];
            }
            dfold = pad ? 0xFFFFFFFFu : din[(size_t)locus * n + i];
            depth[(size_t)locus * lay.n_pad + i] = dfold;
        }
        fold_dmin(dmin, locus, dfold);
        fold_q(qbuf, locus, lay.A, fw);
    }
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

static cudaError_t ingest_prologue(int64_t n_loci, const Layout &lay, const IngestOut &o, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(o.dmin, 0xFF, (size_t)n_loci * 4, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(o.qbuf, 0, (size_t)n_loci * lay.A * 8, s);
    return e;
}
static cudaError_t ingest_epilogue(int64_t n_loci, const Layout &lay, const IngestOut &o, cudaStream_t s) {
    static const bool no_hint = getenv("PG_NOHINT") != nullptr;  // tuning knob: every locus takes the scan's own path
    if (no_hint) return cudaMemsetAsync(o.hint, 0, (size_t)n_loci, s);
    hint_kernel<<<grid_for(n_loci), 256, 0, s>>>(o.qbuf, o.dmin, n_loci, lay.A, lay.n, o.maf, o.one_minus_maf,
                                                 o.min_depth_f, o.hint);
    return cudaGetLastError();
}
template <typename CT>
static cudaError_t launch_ingest_t(const CT *counts, int64_t n_loci, int n, int A_in, int drop_col, const Layout &lay,
                                   const IngestOut &o, cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    cudaError_t e = ingest_prologue(n_loci, lay, o, s);
    if (e != cudaSuccess) return e;
    ingest_counts_kernel<CT><<<grid_for(n_loci * lay.n_pad), 256, 0, s>>>(counts, n_loci, n, A_in, drop_col, lay, o.freq,
                                                                         o.depth, o.dmin, o.qbuf, o.w);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return ingest_epilogue(n_loci, lay, o, s);
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
    cudaError_t e = ingest_prologue(n_loci, lay, o, s);
    if (e != cudaSuccess) return e;
    ingest_freq_kernel<<<grid_for(n_loci * lay.n_pad), 256, 0, s>>>(fin, din, n_loci, n, lay, o.freq, o.depth, o.dmin,
                                                                    o.qbuf, o.w);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return ingest_epilogue(n_loci, lay, o, s);
}
cudaError_t launch_synth(uint64_t seed, int64_t first_locus, int64_t n_loci, int n, int A_in, uint32_t *counts,
                         cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    synth_kernel<<<grid_for(n_loci * n), 256, 0, s>>>(seed, first_locus, n_loci, n, A_in, counts);
    return cudaGetLastError();
}

}  // namespace pg
