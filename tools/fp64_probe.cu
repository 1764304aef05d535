// fp64_probe.cu -- microbenchmark behind the arithmetic design of the scan and kinship kernels: FP64 throughput
// per SM of (a) DFMA with independent chains, (b) mma.sync m8n8k4 f64 (DMMA), (c) mma.sync m16n8k8 f64, as a
// function of the warps per SM and the number of independent accumulators per warp.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/fp64_probe.cu -o gpurun_out/fp64_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                               \
    do {                                                    \
        cudaError_t e_ = (x);                               \
        if (e_ != cudaSuccess) {                            \
            printf("%s: %s\n", #x, cudaGetErrorString(e_)); \
            exit(1);                                        \
        }                                                   \
    } while (0)

template <int CH>
__global__ void k_dfma(double *out, int iters, double a, double b) {
    double v[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += v[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dmma884(double *out, int iters, double a, double b) {
    double d0[CH], d1[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) d0[i] = d1[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma884(d0[i], d1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += d0[i] + d1[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int CH>
__global__ void k_dmma1688(double *out, int iters, double a, double b) {
    double d[CH][4];
    double af[4] = {a, a + 1, a + 2, a + 3}, bf[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < CH; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] = threadIdx.x * 1e-3 + i + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma1688(d[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    printf("%s: %d SMs, %.0f MHz nominal\n", prop.name, sms, khz / 1e3);
    double *out;
    CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    printf("%-12s %5s %3s %12s %14s\n", "kind", "warps", "ch", "TFLOP/s", "FMA/clk/SM@max");
    for (int warps : {4, 8, 16, 32}) {
#define RUN(NAME, KERN, CH, FMA_PER_INST)                                                           \
    {                                                                                               \
        float ms = time_it([&] { KERN<CH><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); });    \
        double fma = (double)sms * warps * (double)iters * CH * (FMA_PER_INST);                     \
        printf("%-12s %5d %3d %12.2f %14.1f\n", NAME, warps, CH, 2 * fma / (ms * 1e-3) / 1e12,      \
               fma / (ms * 1e-3) / sms / (khz * 1e3));                                              \
    }
        RUN("dfma", k_dfma, 8, 32.0)
        RUN("dfma", k_dfma, 16, 32.0)
        RUN("dmma884", k_dmma884, 2, 256.0)
        RUN("dmma884", k_dmma884, 4, 256.0)
        RUN("dmma884", k_dmma884, 8, 256.0)
        RUN("dmma1688", k_dmma1688, 2, 1024.0)
        RUN("dmma1688", k_dmma1688, 4, 1024.0)
    }
    return 0;
}
