// DFMA issue model probe (build + run on the GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_chains tools/probes/dfma_chains.cu
// K independent dependent-chains of three-register DFMAs per thread, W warps per SM sub-partition: cycles per DFMA per
// sub-partition.  Optionally E extra 32-bit selects per DFMA in the stream (are they free?).
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int E>
__global__ void __launch_bounds__(1024) chains(double* out, const double* in, int iters, long long* cyc)
{
    double c[K], a[K], b[K];
    unsigned s[8];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        c[k] = in[threadIdx.x + 32 * k];
        a[k] = in[threadIdx.x + 32 * k + 1000];
        b[k] = in[threadIdx.x + 32 * k + 2000];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = threadIdx.x + e;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                c[k] = fma(a[k], c[k], b[k]);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    // a select that depends on nothing in the FP64 chains
                    asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 7; selp.b32 %0, %1, %2, p; }"
                                 : "=r"(s[(k + e) & 7]) : "r"(s[(k + e + 1) & 7]), "r"(s[(k + e + 2) & 7]), "r"(it));
                }
            }
        }
    }
    const long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) r += c[k];
    unsigned q = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) q ^= s[e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + q;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, int E>
void run(int warps_per_smsp)
{
    const int threads = 128 * warps_per_smsp, blocks = 148, iters = 2000;
    double *out, *in;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaMalloc(&in, sizeof(double) * 4096);
    cudaMemset(in, 0, sizeof(double) * 4096);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    chains<K, E><<<blocks, threads>>>(out, in, 10, cyc);
    chains<K, E><<<blocks, threads>>>(out, in, iters, cyc);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    const double dfma_per_smsp = (double)iters * 8 * K * warps_per_smsp;
    printf("chains/thread %d  warps/SMSP %d  selects/DFMA %d : %.2f cycles per DFMA per SMSP (%.1f%% of the 2-cycle pipe)\n", K,
           warps_per_smsp, E, avg / dfma_per_smsp, 200.0 * dfma_per_smsp / avg);
    cudaFree(out);
    cudaFree(in);
    cudaFree(cyc);
}

int main()
{
    for (int w : {1, 2, 3, 4}) {
        run<1, 0>(w);
        run<2, 0>(w);
        run<3, 0>(w);
        run<4, 0>(w);
        run<6, 0>(w);
        run<8, 0>(w);
    }
    for (int w : {2}) {
        run<2, 1>(w);
        run<2, 2>(w);
        run<4, 1>(w);
        run<4, 2>(w);
        run<8, 1>(w);
        run<8, 2>(w);
    }
    return 0;
}
