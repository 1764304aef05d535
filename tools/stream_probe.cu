// stream_probe.cu -- microbenchmark behind the data-movement design of pg_scan.cuh: how fast can per-warp rings of
// cp.async.bulk copies stream a large array on a B200, as a function of tile size, ring depth, warps per SM and
// the address pattern (each warp walking its own contiguous range vs. all warps interleaved tile by tile)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/stream_probe.cu -o gpurun_out/stream_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("%s: %s\n", #x, cudaGetErrorString(e_));                        \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// pattern 0: warp w owns a contiguous range of `span` tiles at a time (like the scan kernel's groups)
// pattern 1: tile t of the whole grid goes to warp t % total_warps (fully interleaved)
__global__ void probe(const char *__restrict__ src, size_t n_tiles, int tile_bytes, int nbuf, int span, int pattern,
                      int n_copies, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned char *wb = smem + (size_t)warp * ((size_t)nbuf * tile_bytes + 64);
    uint64_t *bars = reinterpret_cast<uint64_t *>(wb);
    unsigned char *bufs = wb + 64;
    if (lane == 0) {
        for (int b = 0; b < nbuf; b++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[b])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t gw = (size_t)blockIdx.x * nwarps + warp, TW = (size_t)gridDim.x * nwarps;
    auto tile_of = [&](size_t k) -> size_t {  // k-th tile of this warp
        if (pattern == 1) return gw + k * TW;
        const size_t blk = k / span, off = k % span;
        return (gw + blk * TW) * span + off;
    };
    size_t mine = 0;
    while (tile_of(mine) < n_tiles) mine++;  // count (cheap enough for a probe)
    auto issue = [&](size_t k, int b) {
        if (lane == 0) {
            const char *s = src + tile_of(k) * (size_t)tile_bytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[b])),
                         "r"(tile_bytes)
                         : "memory");
            const int part = tile_bytes / n_copies;
            for (int c = 0; c < n_copies; c++)
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(bufs + (size_t)b * tile_bytes + (size_t)c * part)),
                    "l"(s + (size_t)c * part), "r"(part), "r"(smem_u32(&bars[b]))
                    : "memory");
        }
    };
    for (int b = 0; b < nbuf && (size_t)b < mine; b++) issue(b, b);
    unsigned long long acc = 0;
    uint32_t phase = 0;
    int b = 0;
    for (size_t k = 0; k < mine; k++) {
        const uint32_t par = (phase >> b) & 1u;
        asm volatile(
            "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra "
            "LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(&bars[b])),
            "r"(par)
            : "memory");
        phase ^= 1u << b;
        acc += *reinterpret_cast<const unsigned long long *>(bufs + (size_t)b * tile_bytes + lane * 8);
        __syncwarp();
        if (k + nbuf < mine) issue(k + nbuf, b);
        b = (b + 1 == nbuf) ? 0 : b + 1;
    }
    if (acc == 0x1234567ull) sink[0] = acc;
}

// plain coalesced 128-bit loads, 4 in flight per thread (the STREAM-style ceiling)
__global__ void probe_ldg(const uint4 *__restrict__ src, size_t n16, unsigned long long *sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride),
              d = __ldg(src + i + 3 * stride);
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 0x1234567ull) sink[0] = acc;
}

int main(int argc, char **argv) {
    const size_t bytes = (size_t)16 << 30;
    char *d;
    unsigned long long *sink;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(d, 1, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            probe_ldg<<<148 * 8, 512>>>((const uint4 *)d, bytes / 16, sink);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("ldg128 x4            : %7.1f GB/s\n", bytes / ms / 1e6);
        }
    }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int tiles[] = {1024, 2048, 4096, 8192, 16384};
    const int warps[] = {4, 8, 12, 16};
    const int nbufs[] = {2, 3, 4, 6};
    printf("%6s %5s %5s %4s %4s %5s %9s\n", "tile", "warps", "nbuf", "pat", "ncp", "KB/SM", "GB/s");
    for (int pattern = 0; pattern < 2; pattern++)
        for (int tb : tiles)
            for (int w : warps)
                for (int nb : nbufs)
                    for (int ncp : {1, 2}) {
                        const size_t smem = (size_t)w * ((size_t)nb * tb + 64);
                        if (smem > 220 * 1024) continue;
                        if (ncp == 2 && !(tb == 4096 && nb == 3)) continue;
                        const size_t n_tiles = bytes / tb;
                        const int span = (252 * 1024) / tb;  // ~ one group of 7 loci
                        float ms = 0;
                        for (int rep = 0; rep < 2; rep++) {
                            cudaEventRecord(e0);
                            probe<<<148, w * 32, smem>>>(d, n_tiles, tb, nb, span, pattern, ncp, sink);
                            cudaEventRecord(e1);
                            CK(cudaEventSynchronize(e1));
                            cudaEventElapsedTime(&ms, e0, e1);
                        }
                        printf("%6d %5d %5d %4d %4d %5.0f %9.1f\n", tb, w, nb, pattern, ncp, smem / 1024.0,
                               bytes / ms / 1e6);
                    }
    return 0;
}
