// microbench.cuh -- small device probes behind kw_fd1d_fp64_peak() / kw_fd1d_microbench():
// the measured FP64-FMA roofline denominator and the latencies DESIGN.md's model uses.
#pragma once
#include <cuda_runtime.h>

namespace kwfd1d {

// 8 independent DFMA chains per thread: throughput-bound on the FP64 pipe.
__global__ void __launch_bounds__(1024) dfma_throughput_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
            a4 = fma(a4, m, c);
            a5 = fma(a5, m, c);
            a6 = fma(a6, m, c);
            a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

// cycles per op of a few dependent chains, one warp (or one 128-thread CTA for the barrier)
// res: 0 DFMA, 1 shfl_up(double), 2 shfl_up(double)+DFMA (one scan level), 3 __syncthreads (4 warps),
//      4 LDS.64 dependent, 5 DMNMX(fmax) dependent
__global__ void latency_kernel(double* res, double seed, int iters)
{
    __shared__ double sm[256];
    const int lane = threadIdx.x & 31;
    sm[threadIdx.x] = (double)((threadIdx.x * 7 + 1) & 255);
    sm[threadIdx.x + 128] = (double)((threadIdx.x * 13 + 5) & 255);
    __syncthreads();
    double a = seed + lane;
    const double m = 0.999999, c = 1e-9;
    long long t0, t1;

    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(a, m, c);
    }
    t1 = clock64();
    if (threadIdx.x == 0) res[0] = (double)(t1 - t0) / (16.0 * iters);

    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = __shfl_up_sync(0xffffffffu, a, 1);
    }
    t1 = clock64();
    if (threadIdx.x == 0) res[1] = (double)(t1 - t0) / (16.0 * iters);

    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(m, __shfl_up_sync(0xffffffffu, a, 1), a);
    }
    t1 = clock64();
    if (threadIdx.x == 0) res[2] = (double)(t1 - t0) / (16.0 * iters);

    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) __syncthreads();
    }
    t1 = clock64();
    if (threadIdx.x == 0) res[3] = (double)(t1 - t0) / (16.0 * iters);

    int idx = threadIdx.x;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) idx = (int)sm[idx & 255];
    }
    t1 = clock64();
    if (threadIdx.x == 0) res[4] = (double)(t1 - t0) / (16.0 * iters);

    double b = a * 0.5;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) b = fmax(b + c, a);  // DADD + DMNMX dependent pair
    }
    t1 = clock64();
    if (threadIdx.x == 0) res[5] = (double)(t1 - t0) / (16.0 * iters);

    if (a + b + idx == 12345.678) res[7] = a;
}

}  // namespace kwfd1d
