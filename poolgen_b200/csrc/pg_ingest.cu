// pg_ingest.cu -- ingestion kernels: parsed counts (or a host-built frequency matrix) -> the chunked
// f64 first-stage frequency matrix + depth vector the scan kernel streams, and the synthetic workload
// generator.  Frequencies are one IEEE division of exactly representable integers, i.e. bit-identical
// to LocusCounts::to_frequencies (src/base/sync.rs:166-192).  Every ingest also leaves dmin[locus] = the smallest
// pool depth of the locus (the caller presets dmin to 0xFFFFFFFF).
#include "pg_internal.h"

namespace pg {

// lanes of the warp that hold the same locus fold their depths, one atomic per (warp, locus)
__device__ __forceinline__ void fold_dmin(uint32_t *dmin, int64_t locus, uint32_t d) {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, (unsigned long long)locus);
    const unsigned m = __reduce_min_sync(peers, d);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicMin(dmin + locus, m);
}

template <typename CT>
__global__ void __launch_bounds__(256) ingest_counts_kernel(const CT *__restrict__ counts, int64_t n_loci, int n,
                                                            int A_in, int drop_col, Layout lay,
                                                            double *__restrict__ freq, uint32_t *__restrict__ depth,
                                                            uint32_t *__restrict__ dmin) {
    const int64_t total = n_loci * lay.n_pad;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t locus = idx / lay.n_pad;
        const int i = (int)(idx - locus * lay.n_pad);
        double *fl = freq + (size_t)locus * lay.freq_stride();
        if (i >= n) {
            for (int j = 0; j < lay.A; j++) fl[lay.freq_off(i, j)] = 0.0;
            depth[(size_t)locus * lay.n_pad + i] = 0xFFFFFFFFu;
            fold_dmin(dmin, locus, 0xFFFFFFFFu);
            continue;
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
        const double dd = (double)d;
        for (int j = 0; j < lay.A; j++) fl[lay.freq_off(i, j)] = (d == 0) ? nan("") : (double)c[j] / dd;
        depth[(size_t)locus * lay.n_pad + i] = (uint32_t)d;
        fold_dmin(dmin, locus, (uint32_t)d);
    }
}

__global__ void __launch_bounds__(256) ingest_freq_kernel(const double *__restrict__ fin, const uint32_t *__restrict__ din,
                                                          int64_t n_loci, int n, Layout lay,
                                                          double *__restrict__ freq, uint32_t *__restrict__ depth,
                                                          uint32_t *__restrict__ dmin) {
    const int64_t total = n_loci * lay.n_pad;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t locus = idx / lay.n_pad;
        const int i = (int)(idx - locus * lay.n_pad);
        double *fl = freq + (size_t)locus * lay.freq_stride();
        const bool pad = i >= n;
        for (int j = 0; j < lay.A; j++)
            fl[lay.freq_off(i, j)] = pad ? 0.0 : fin[((size_t)locus * lay.A + j) * n + i];
        const uint32_t d = pad ? 0xFFFFFFFFu : din[(size_t)locus * n + i];
        depth[(size_t)locus * lay.n_pad + i] = d;
        fold_dmin(dmin, locus, d);
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

cudaError_t launch_ingest_u32(const uint32_t *counts, int64_t n_loci, int n, int A_in, int drop_col,
                              const Layout &lay, double *freq, uint32_t *depth, uint32_t *dmin, cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dmin, 0xFF, (size_t)n_loci * 4, s);
    if (e != cudaSuccess) return e;
    ingest_counts_kernel<uint32_t><<<grid_for(n_loci * lay.n_pad), 256, 0, s>>>(counts, n_loci, n, A_in, drop_col,
                                                                               lay, freq, depth, dmin);
    return cudaGetLastError();
}
cudaError_t launch_ingest_u16(const uint16_t *counts, int64_t n_loci, int n, int A_in, int drop_col,
                              const Layout &lay, double *freq, uint32_t *depth, uint32_t *dmin, cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dmin, 0xFF, (size_t)n_loci * 4, s);
    if (e != cudaSuccess) return e;
    ingest_counts_kernel<uint16_t><<<grid_for(n_loci * lay.n_pad), 256, 0, s>>>(counts, n_loci, n, A_in, drop_col,
                                                                               lay, freq, depth, dmin);
    return cudaGetLastError();
}
cudaError_t launch_ingest_u8(const uint8_t *counts, int64_t n_loci, int n, int A_in, int drop_col,
                             const Layout &lay, double *freq, uint32_t *depth, uint32_t *dmin, cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dmin, 0xFF, (size_t)n_loci * 4, s);
    if (e != cudaSuccess) return e;
    ingest_counts_kernel<uint8_t><<<grid_for(n_loci * lay.n_pad), 256, 0, s>>>(counts, n_loci, n, A_in, drop_col,
                                                                              lay, freq, depth, dmin);
    return cudaGetLastError();
}
cudaError_t launch_ingest_freq(const double *fin, const uint32_t *din, int64_t n_loci, int n, const Layout &lay,
                               double *freq, uint32_t *depth, uint32_t *dmin, cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dmin, 0xFF, (size_t)n_loci * 4, s);
    if (e != cudaSuccess) return e;
    ingest_freq_kernel<<<grid_for(n_loci * lay.n_pad), 256, 0, s>>>(fin, din, n_loci, n, lay, freq, depth, dmin);
    return cudaGetLastError();
}
cudaError_t launch_synth(uint64_t seed, int64_t first_locus, int64_t n_loci, int n, int A_in, uint32_t *counts,
                         cudaStream_t s) {
    if (n_loci <= 0) return cudaSuccess;
    synth_kernel<<<grid_for(n_loci * n), 256, 0, s>>>(seed, first_locus, n_loci, n, A_in, counts);
    return cudaGetLastError();
}

}  // namespace pg
