// microbench.cuh -- small device probes behind kw_fd1d_fp64_peak() / kw_fd1d_microbench():
// the measured FP64-FMA roofline denominator and the latencies DESIGN.md's model uses.
#pragma once
#include <cuda_runtime.h>

#include "tmem.cuh"

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

namespace kwfd1d {
// ---- Tensor-memory probes (kw_fd1d_tmem_probe): TMEM as a per-thread constant store ---------
// Every CTA of 128 threads allocates 128 TMEM columns (64 doubles per thread), four CTAs per SM.
// One "round" reads all 64 doubles back in eight tcgen05.ld.32x32b.x16 and optionally feeds them
// to DFMAs.  MODE 0: DFMA only (16 per chunk); 1: ld+wait only; 2: ld+wait+8 DFMA;
// 3: ld+wait+16 DFMA; 4: one x32 ld per two chunks + wait; 5: ld issued one chunk ahead + 16 DFMA.

template <int MODE>
__global__ void __launch_bounds__(128, 4) tmem_probe_kernel(double* out, long long* cycles, int iters, double seed)
{
    __shared__ uint32_t s_taddr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem::alloc128((uint32_t)__cvta_generic_to_shared(&s_taddr));
    tmem::fence_before();
    __syncthreads();
    tmem::fence_after();
    const uint32_t base = s_taddr + ((uint32_t)(warp & 3) << 21);  // lane field = 32 * (warp % 4), bits 31:16
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = seed + threadIdx.x + i;
    // fill: chunk c holds 1 - (c*8+i) * 2^-40 + tid * 2^-30
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = 1.0 - (double)(c * 8 + i) * 0x1p-40 + (double)threadIdx.x * 0x1p-30;
        tmem::st8(base + c * 16, d);
    }
    tmem::wait_st();
    const double cc = 1e-9;
    double chk = 0.;
    uint32_t xr = 0;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], 0.999999, cc);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], 0.999999, cc);
            }
        } else if (MODE == 5) {
            double d[8], e[8];
            tmem::ld8(base, d);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                tmem::wait_ld_dep(d);
                if (c < 7) tmem::ld8(base + (c + 1) * 16, e);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], d[i], cc);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], d[i], cc);
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] = e[i];
            }
        } else if (MODE == 1 || MODE == 4) {
#pragma unroll
            for (int c = 0; c < 8; ++c) xr ^= tmem::ld16_raw(base + c * 16);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double d[8];
                tmem::ld8(base + c * 16, d);
                tmem::wait_ld_dep(d);

                if (MODE >= 2 && MODE != 4) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], d[i], cc);
                }
                if (MODE == 3) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], d[i], cc);
                }
            }
        }
    }
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    double s = chk + (double)xr;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (threadIdx.x == 0) {
        // per-CTA record: SM id, start and end of the timed loop on that SM's clock
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        cycles[2 + 3 * blockIdx.x] = (long long)smid;
        cycles[3 + 3 * blockIdx.x] = t0;
        cycles[4 + 3 * blockIdx.x] = t1;
        if (blockIdx.x == 0) {
            cycles[0] = t1 - t0;
            cycles[2 + 3 * gridDim.x] = (long long)(g1 - g0);  // the same loop in ns (globaltimer)
        }
    }
    if (s == 12345.678) out[0] = s;
    // read-back check of one value: chunk 3, element 5
    {
        double d[8];
        tmem::ld8(base + 3 * 16, d);
        tmem::wait_ld_dep(d);
        const double want = 1.0 - (double)(3 * 8 + 5) * 0x1p-40 + (double)threadIdx.x * 0x1p-30;
        if (d[5] != want) atomicAdd((unsigned long long*)&cycles[1], 1ull);
    }
    tmem::fence_before();
    __syncthreads();
    if (warp == 0) tmem::dealloc128(s_taddr);
}

}  // namespace kwfd1d

namespace kwfd1d {
// ---- DFMA issue rate against the number of distinct REGISTER source operands ----------------
// KIND 1: acc = fma(acc, m, c) (one register source); 2: acc = fma(x_i, c, acc) (two);
// 3: acc = fma(x_i, y_i, acc) (three distinct); 4: acc = fma(x_i, Y, acc), Y shared by all eight
// chains (operand-reuse candidate).  Eight independent chains per thread.
template <int KIND>
__global__ void __launch_bounds__(128) dfma_operand_kernel(const double* in, double* out, int iters)
{
    double x[8], y[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = in[(threadIdx.x + i) & 63];
        y[i] = in[(threadIdx.x + 8 + i) & 63];
        acc[i] = in[(threadIdx.x + 16 + i) & 63];
    }
    const double Y = in[blockIdx.x & 63];
    int s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s0[i] = (int)threadIdx.x + i;
        s1[i] = (int)blockIdx.x - i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 1) acc[i] = fma(acc[i], 0.999999, 1e-9);
                if (KIND == 2) acc[i] = fma(x[i], 1e-9, acc[i]);
                if (KIND == 3) acc[i] = fma(x[i], y[i], acc[i]);
                if (KIND == 4) acc[i] = fma(x[i], Y, acc[i]);
                if (KIND == 5 || KIND == 6) {
                    // three-source DFMA plus one (5) or two (6) 32-bit integer ops on unrelated registers:
                    // do non-FP64 instructions that read registers slow the FP64 stream down?
                    acc[i] = fma(x[i], y[i], acc[i]);
                    s0[i] = s0[i] + s1[(i + 1) & 7];          // IADD3, two register sources
                    if (KIND == 6) s1[i] = s1[i] ^ s0[(i + 3) & 7];  // LOP3, two register sources
                }
            }
        }
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i] + (double)(s0[i] ^ s1[i]);
    if (s == 12345.678) out[0] = s;
}
}  // namespace kwfd1d
