// fd1d_common.cuh -- shared device helpers of the Fd1d kernels (sm_100a).
//
// The scheme is the reference's (src/Math/kwFd1d.cpp:61-136): theta = 1/2 Crank-Nicolson,
// B = 1 - dt/2 A assembled on the non-uniform sinh grid with one-sided, diffusion-free
// boundary rows, Thomas solve, explicit projection v = max(v, payoff) on nodes 0..xDim-2.
//
// What is hoisted (DESIGN.md "Algebra"): dt is constant (t/(tDim-1)), so B and its LU are
// time-invariant and C = 2I - B, hence one step is  v <- max(2 B^-1 v - v, payoff).
// With beta_j the Thomas pivots and pivot-scaled unknowns y~ = beta*y, u~ = beta*u:
//     y~_j = v_j + a~_j y~_{j-1},   a~_j = -bl_j / beta_{j-1}
//     u~_j = y~_j + g~_j u~_{j+1},  g~_j = -bu_j / beta_{j+1}
//     v'_j = max(D_j u~_j - v_j, p_j),   D_j = 2 / beta_j
// i.e. 4 FP64-pipe instructions per node-step serial (3 FMA + 1 max), no division.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/kw_fd1d.h"

namespace kwfd1d {

// One batch of PDEs as the kernels see it.  With chain compression several options share
// a PDE (reference src/Pricer/kwFd1d.cpp:28-65): pde_rep[p] is the option whose (t,r,q,z,e,w)
// define PDE p and csr_start/csr_opt list the options priced from it.  All three null =
// identity (PDE p == option p).
struct Fd1dBatch {
    const kw_option* opts;
    const uint32_t* pde_rep;
    const uint32_t* csr_start;
    const uint32_t* csr_opt;
    const uint32_t* csr_cnt;    // device-side compression (compress.cuh): members per PDE; then
                                // csr_start[p] is the segment start and there is no csr_start[m]
    const uint32_t* n_pde_dev;  // device-side compression: the PDE count lives in HBM (overrides n_pde)
    double* prices;
    unsigned int* status;  // [0] = number of out-of-range options, [1] = smallest such index,
                           // [2..7] = PDEs marched per carry mode / scan-level count, [8] = work counter (fd1d_iw.cuh)
    uint32_t n_pde;
    uint32_t pde_base;     // first PDE of this launch (layout A chunks the batch)
    int32_t tDim;
    int32_t xDim;
    int32_t max_mode;      // layout B: highest carry-truncation mode allowed (4 = all, 0 = exact)
    double density;
    double scale;
    double* prices_eu;     // fused FD1D-BS march only (fd1d_warp_bs.cuh): the European solution's prices
    // Values the kernels must not be able to fold (fd1d_warp.cuh, SPLIT): opq_lim[] = INT32_MAX ("step < lim" is
    // always true), opq_zero = 0.  capi.cu: make_batch().
    int32_t opq_lim[4];
    uint32_t opq_zero;
    // fd1d_iw.cuh: the next PDE to hand out (status[8], zeroed by status_reset_kernel before every launch)
    unsigned int* work_counter;
    // PDE-count class of this launch: with the count known only on the device (n_pde_dev), capi.cu launches the kernel of
    // every class and the ones whose class [pde_lo, pde_hi) does not hold the count see an empty batch
    uint32_t pde_lo, pde_hi;
    // Long chains (fd1d_iw.cuh): a warp that finds more than KW_LONG_CHAIN options on its PDE does not interpolate them itself;
    // it leaves the final v in a workspace slot and fd1d_long_value_kernel prices them with a CTA per chain afterwards.
    double* long_ws;         // [long_cap][nodes per PDE tile], or null (feature off: no chain compression)
    uint32_t* long_meta;     // [long_cap] PDE id | (European copy of the fused FD1D-BS march) << 31
    unsigned int* long_count;  // slots handed out (status[9], zeroed before every launch)
    uint32_t long_cap;
    uint32_t n_opt;          // options in the batch
};
#define KW_LONG_CHAIN 512u

__device__ __forceinline__ uint32_t batch_n_pde(const Fd1dBatch& B)
{
    const uint32_t n = B.n_pde_dev ? __ldg(B.n_pde_dev) : B.n_pde;
    return (n >= B.pde_lo && n < B.pde_hi) ? n : 0u;
}

// [q0, q1) = this PDE's entries of csr_opt
__device__ __forceinline__ void chain_range(const Fd1dBatch& B, uint32_t pde, uint32_t& q0, uint32_t& q1)
{
    if (!B.csr_start) {
        q0 = pde;
        q1 = pde + 1;
    } else if (B.csr_cnt) {
        q0 = __ldg(B.csr_start + pde);
        q1 = q0 + __ldg(B.csr_cnt + pde);
    } else {
        q0 = __ldg(B.csr_start + pde);
        q1 = __ldg(B.csr_start + pde + 1);
    }
}

struct PdeScalars {
    double a0, ax, axx;  // src/Pricer/kwFd1d.cpp:80-83
    double hdt;          // theta*dt with theta = 0.5 (src/Math/kwFd1d.h:31)
    double yMin, yMax, dy;
    bool american, put;
};

__device__ __forceinline__ kw_option load_option(const kw_option* p)
{
    // 56-byte struct, 8-byte aligned: seven 64-bit loads through the read-only path
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p);
    kw_option o;
    o.t = __longlong_as_double(__ldg(q + 0));
    o.k = __longlong_as_double(__ldg(q + 1));
    o.z = __longlong_as_double(__ldg(q + 2));
    o.r = __longlong_as_double(__ldg(q + 3));
    o.q = __longlong_as_double(__ldg(q + 4));
    o.s = __longlong_as_double(__ldg(q + 5));
    const unsigned long long ew = __ldg(q + 6);
    o.e = (uint8_t)(ew & 0xff);
    o.w = (int8_t)((ew >> 8) & 0xff);
    return o;
}

// src/Pricer/kwFd1d.cpp:68-86 (coefficients), :95-99 (dt), :105-119 (y range).  The
// _rn intrinsics keep the reference's evaluation order free of FMA contraction.
__device__ __forceinline__ PdeScalars pde_scalars(const kw_option& o, const Fd1dBatch& B)
{
    PdeScalars s;
    const double zz = __dmul_rn(o.z, o.z);
    s.a0 = -o.r;
    s.ax = __dsub_rn(__dsub_rn(o.r, o.q), zz / 2);
    s.axx = zz / 2;
    const double dt = o.t / (double)(B.tDim - 1);
    s.hdt = 0.5 * dt;
    const double half = __dmul_rn(__dmul_rn(B.scale, o.z), sqrt(o.t));
    s.yMin = asinh((0. - half) / B.density);
    s.yMax = asinh((0. + half) / B.density);
    s.dy = 1. / (double)(B.xDim - 1);
    s.american = o.e != 0;
    s.put = o.w < 0;
    return s;
}

// x_j, src/Pricer/kwFd1d.cpp:120-124
__device__ __forceinline__ double x_node(const PdeScalars& s, double density, int j)
{
    const double yj = __dmul_rn((double)j, s.dy);
    const double arg = __dadd_rn(__dmul_rn(s.yMin, 1.0 - yj), __dmul_rn(s.yMax, yj));
    return __dmul_rn(density, sinh(arg));
}

// payoff in strike units, src/Pricer/kwFd1d.cpp:127-139
__device__ __forceinline__ double payoff_node(bool put, double x)
{
    const double ex = exp(x);
    const double p = put ? (1. - ex) : (ex - 1.);
    return 0 < p ? p : 0.;
}

// Row j of B = 1 - theta*dt*A, src/Math/kwFd1d.cpp:73-114, with the constant dt.
// xm, x0, xp = x_{j-1}, x_j, x_{j+1} (unused neighbours may be anything).
__device__ __forceinline__ void b_row(const PdeScalars& s, int j, int xDim, double xm, double x0,
                                      double xp, double& bl, double& b, double& bu)
{
    const double hdt = s.hdt, a0 = s.a0, ax = s.ax, axx = s.axx;
    if (j >= xDim) {  // padding rows: identity, decoupled
        bl = 0.;
        b = 1.;
        bu = 0.;
    } else if (j == 0) {
        const double inv_dx = 1. / (xp - x0);
        bl = 0.;
        b = __dsub_rn(1., __dmul_rn(hdt, __dsub_rn(a0, __dmul_rn(inv_dx, ax))));
        bu = -__dmul_rn(hdt, __dmul_rn(inv_dx, ax));
    } else if (j == xDim - 1) {
        const double inv_dx = 1. / (x0 - xm);
        bl = -__dmul_rn(hdt, __dmul_rn(-inv_dx, ax));
        b = __dsub_rn(1., __dmul_rn(hdt, __dadd_rn(a0, __dmul_rn(inv_dx, ax))));
        bu = 0.;
    } else {
        const double inv_dxu = 1. / (xp - x0);
        const double inv_dxm = 1. / (xp - xm);
        const double inv_dxd = 1. / (x0 - xm);
        const double inv_dx2u = __dmul_rn(__dmul_rn(2., inv_dxu), inv_dxm);
        const double inv_dx2m = __dmul_rn(__dmul_rn(2., inv_dxd), inv_dxu);
        const double inv_dx2l = __dmul_rn(__dmul_rn(2., inv_dxd), inv_dxm);
        bl = -__dmul_rn(hdt, __dadd_rn(__dmul_rn(-inv_dxm, ax), __dmul_rn(inv_dx2l, axx)));
        b = __dsub_rn(1., __dmul_rn(hdt, __dsub_rn(a0, __dmul_rn(inv_dx2m, axx))));
        bu = -__dmul_rn(hdt, __dadd_rn(__dmul_rn(inv_dxm, ax), __dmul_rn(inv_dx2u, axx)));
    }
}

// The same row with the spacing inverses shared between neighbouring rows: 1 / (x_j - x_{j-1}) is row j's inv_dxd and row
// j-1's inv_dxu -- the same operands, the same IEEE quotient -- so a sweep over consecutive rows needs two divisions per row
// instead of three.  inv_d: in, 1 / (x0 - xm) (unused for j = 0); inv_u: out, 1 / (xp - x0) (undefined for j >= xDim - 1).
__device__ __forceinline__ void b_row_chained(const PdeScalars& s, int j, int xDim, double xm, double x0, double xp,
                                              double inv_d, double& inv_u, double& bl, double& b, double& bu)
{
    const double hdt = s.hdt, a0 = s.a0, ax = s.ax, axx = s.axx;
    if (j >= xDim) {  // padding rows: identity, decoupled
        bl = 0.;
        b = 1.;
        bu = 0.;
    } else if (j == 0) {
        const double inv_dx = 1. / (xp - x0);
        inv_u = inv_dx;
        bl = 0.;
        b = __dsub_rn(1., __dmul_rn(hdt, __dsub_rn(a0, __dmul_rn(inv_dx, ax))));
        bu = -__dmul_rn(hdt, __dmul_rn(inv_dx, ax));
    } else if (j == xDim - 1) {
        const double inv_dx = inv_d;
        bl = -__dmul_rn(hdt, __dmul_rn(-inv_dx, ax));
        b = __dsub_rn(1., __dmul_rn(hdt, __dadd_rn(a0, __dmul_rn(inv_dx, ax))));
        bu = 0.;
    } else {
        const double inv_dxu = 1. / (xp - x0);
        const double inv_dxm = 1. / (xp - xm);
        const double inv_dxd = inv_d;
        inv_u = inv_dxu;
        const double inv_dx2u = __dmul_rn(__dmul_rn(2., inv_dxu), inv_dxm);
        const double inv_dx2m = __dmul_rn(__dmul_rn(2., inv_dxd), inv_dxu);
        const double inv_dx2l = __dmul_rn(__dmul_rn(2., inv_dxd), inv_dxm);
        bl = -__dmul_rn(hdt, __dadd_rn(__dmul_rn(-inv_dxm, ax), __dmul_rn(inv_dx2l, axx)));
        b = __dsub_rn(1., __dmul_rn(hdt, __dsub_rn(a0, __dmul_rn(inv_dx2m, axx))));
        bu = -__dmul_rn(hdt, __dadd_rn(__dmul_rn(inv_dxm, ax), __dmul_rn(inv_dx2u, axx)));
    }
}

// Fd1d::value, src/Math/kwFd1d.cpp:139-158 (+ the k multiplication of
// src/Pricer/kwFd1d.cpp:156).  x is increasing, so the reference's linear search for the first
// x[xi] >= x_ is a lower_bound.  XS / VS are callables j -> x_j / v_j.
template <class XS, class VS>
__device__ __forceinline__ void price_option(const Fd1dBatch& B, uint32_t oi, XS xs, VS vs)
{
    const kw_option o = load_option(B.opts + oi);
    const double xq = log(o.s / o.k);
    int lo = 0, hi = B.xDim;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xs(mid) < xq)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo == 0 || lo == B.xDim) {
        B.prices[oi] = CUDART_NAN;
        atomicAdd(&B.status[0], 1u);
        atomicMin(&B.status[1], oi);
        return;
    }
    const double x1 = xs(lo), x0 = xs(lo - 1);
    const double num = __dadd_rn(__dmul_rn(x1 - xq, vs(lo - 1)), __dmul_rn(xq - x0, vs(lo)));
    B.prices[oi] = __dmul_rn(o.k, num / (x1 - x0));
}

// The same for a caller that does not keep the grid (fd1d_iw.cuh recomputes x_j = density sinh(...), ~100 instructions a node):
// the position of x_ on the sinh grid is known analytically, so the lower bound is found by stepping from ceil of the fractional
// index until x[lo-1] < x_ <= x[lo] holds for the COMPUTED nodes -- the answer of the reference's search, with 2-3 grid evaluations
// instead of log2(xDim) + 2.
template <class VS>
__device__ __forceinline__ void price_option_sinh_grid(const Fd1dBatch& B, const PdeScalars& sc, uint32_t oi, VS vs)
{
    const kw_option o = load_option(B.opts + oi);
    const double xq = log(o.s / o.k);
    const int xDim = B.xDim;
    const double t = (asinh(xq / B.density) - sc.yMin) / (sc.yMax - sc.yMin) * (double)(xDim - 1);
    int lo = !(t > 0.) ? 0 : (t >= (double)xDim ? xDim : (int)ceil(t));  // (NaN: 0, as the search below then answers)
    double x_hi = lo < xDim ? x_node(sc, B.density, lo) : CUDART_INF;      // x[lo]
    double x_lw = lo > 0 ? x_node(sc, B.density, lo - 1) : -CUDART_INF;    // x[lo - 1]
    while (lo > 0 && !(x_lw < xq)) {
        --lo;
        x_hi = x_lw;
        x_lw = lo > 0 ? x_node(sc, B.density, lo - 1) : -CUDART_INF;
    }
    while (lo < xDim && x_hi < xq) {
        ++lo;
        x_lw = x_hi;
        x_hi = lo < xDim ? x_node(sc, B.density, lo) : CUDART_INF;
    }
    if (lo == 0 || lo == xDim) {
        B.prices[oi] = CUDART_NAN;
        atomicAdd(&B.status[0], 1u);
        atomicMin(&B.status[1], oi);
        return;
    }
    const double x1 = x_hi, x0 = x_lw;
    const double num = __dadd_rn(__dmul_rn(x1 - xq, vs(lo - 1)), __dmul_rn(xq - x0, vs(lo)));
    B.prices[oi] = __dmul_rn(o.k, num / (x1 - x0));
}

// The options of the chains that fd1d_iw_kernel deferred (Fd1dBatch::long_ws): a thread per option, gridDim / chains CTAs per chain.
__global__ void __launch_bounds__(256) fd1d_long_value_kernel(const Fd1dBatch B, int XT)
{
    const uint32_t cnt = min(*B.long_count, B.long_cap);
    if (cnt == 0) return;
    const uint32_t seg_n = max(1u, gridDim.x / cnt);  // CTAs per chain: few long chains spread over the whole grid
    for (uint32_t b = blockIdx.x; b < cnt * seg_n; b += gridDim.x) {
        const uint32_t s = b % cnt, seg = b / cnt;
        const uint32_t meta = B.long_meta[s];
        const uint32_t pde = meta & 0x7fffffffu;
        const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + pde) : pde;
        const PdeScalars sc = pde_scalars(load_option(B.opts + rep), B);
        Fd1dBatch Bo = B;
        if (meta >> 31) Bo.prices = B.prices_eu;
        uint32_t q0, q1;
        chain_range(B, pde, q0, q1);
        const double* v = B.long_ws + (size_t)s * XT;
        for (uint32_t q = q0 + seg * blockDim.x + threadIdx.x; q < q1; q += seg_n * blockDim.x) {
            const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
            price_option_sinh_grid(Bo, sc, oi, [&](int j) { return __ldg(v + j); });
        }
    }
}

}  // namespace kwfd1d
