// fd1d_soa.cuh -- Layout A: one THREAD per PDE, Thomas sweeps over a batch-interleaved,
// coalesced structure-of-arrays in global memory: element j of PDE i lives at buf[j*n + i].
//
// This is the HBM-bound comparison layout the north star asks for (and the fall-back for
// x grids too large for Layout B's register tiles): per node-step it streams
// a~, v (read) + y~ (write) forward and g~, y~, D, v, p (read) + v (write) backward = 72 B.
// Same arithmetic as Layout B without the partition: exactly the serial recurrences of
// fd1d_common.cuh.  Replaces the same reference functions (src/Math/kwFd1d.cpp:11-158,
// src/Math/kwMath.cpp:16-49).
#pragma once
#include "fd1d_common.cuh"

namespace kwfd1d {

struct SoaWork {
    double* A;   // a~   [xDim][n]
    double* G;   // g~
    double* D;   // 2/beta
    double* PR;  // projection floor (payoff or -inf)
    double* V;   // solution
    double* Y;   // forward-sweep scratch
    double* X;   // x grid (epilogue only)
    uint32_t n;  // PDEs in this chunk (row pitch)
};

// grid/payoff/LU set-up, serial in x per thread (pivots exactly in the reference's order,
// src/Math/kwMath.cpp:28-41, with the hoisted constant dt)
__global__ void __launch_bounds__(128) fd1d_soa_setup_kernel(const Fd1dBatch B, const SoaWork W)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n) return;
    const uint32_t pde = B.pde_base + i;
    const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + pde) : pde;
    const kw_option opt = load_option(B.opts + rep);
    const PdeScalars sc = pde_scalars(opt, B);
    const int xDim = B.xDim;
    const size_t n = W.n;

    double xm = 0., x0 = x_node(sc, B.density, 0), xp = x_node(sc, B.density, 1);
    double beta_prev = CUDART_INF, bu_prev = 0., ib_prev = 0.;
    for (int j = 0; j < xDim; ++j) {
        double bl, b, bu;
        b_row(sc, j, xDim, xm, x0, xp, bl, b, bu);
        const double gam = bu_prev / beta_prev;
        const double beta = __dsub_rn(b, __dmul_rn(bl, gam));
        const double ib = 1. / beta;
        const double p = payoff_node(sc.put, x0);
        const size_t o = (size_t)j * n + i;
        W.X[o] = x0;
        W.V[o] = p;
        W.PR[o] = (sc.american && j < xDim - 1) ? p : -CUDART_INF;
        W.A[o] = -bl * ib_prev;
        W.D[o] = 2. * ib;
        if (j > 0) W.G[o - n] = -bu_prev * ib;
        if (j == xDim - 1) W.G[o] = 0.;
        beta_prev = beta;
        bu_prev = bu;
        ib_prev = ib;
        xm = x0;
        x0 = xp;
        xp = x_node(sc, B.density, j + 2);
    }
}

// the whole time march, one thread per PDE
__global__ void __launch_bounds__(128) fd1d_soa_march_kernel(const Fd1dBatch B, const SoaWork W)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n) return;
    const int xDim = B.xDim;
    const size_t n = W.n;
    const double* __restrict__ A = W.A + i;
    const double* __restrict__ G = W.G + i;
    const double* __restrict__ D = W.D + i;
    const double* __restrict__ PR = W.PR + i;
    double* __restrict__ V = W.V + i;
    double* __restrict__ Y = W.Y + i;

    for (int step = 0; step < B.tDim - 1; ++step) {
        double y = 0.;
#pragma unroll 8
        for (int j = 0; j < xDim; ++j) {
            const size_t o = (size_t)j * n;
            y = fma(A[o], y, V[o]);
            Y[o] = y;
        }
        double u = 0.;
#pragma unroll 8
        for (int j = xDim - 1; j >= 0; --j) {
            const size_t o = (size_t)j * n;
            u = fma(G[o], u, Y[o]);
            const double r = fma(D[o], u, -V[o]);
            V[o] = fmax(r, PR[o]);
        }
    }
}

// Fd1d::value for every option of the chunk's chains, one thread per PDE
__global__ void __launch_bounds__(128) fd1d_soa_value_kernel(const Fd1dBatch B, const SoaWork W)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n) return;
    const uint32_t pde = B.pde_base + i;
    const size_t n = W.n;
    const double* X = W.X + i;
    const double* V = W.V + i;
    uint32_t q0, q1;
    chain_range(B, pde, q0, q1);
    for (uint32_t q = q0; q < q1; ++q) {
        const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
        price_option(B, oi, [&](int j) { return X[(size_t)j * n]; }, [&](int j) { return V[(size_t)j * n]; });
    }
}

}  // namespace kwfd1d
