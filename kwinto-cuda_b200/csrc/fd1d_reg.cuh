// fd1d_reg.cuh -- Layout B: one CTA per PDE, the whole time march in one launch, x-grid
// coefficients in REGISTERS (optionally the payoff and one fix-up array in shared memory).
//
// Replaces Fd1d::solve/solveOne + solveTridiagonal + Fd1d::value of the reference
// (src/Math/kwFd1d.cpp:11-158, src/Math/kwMath.cpp:16-49) and the grid/payoff set-up of
// Fd1d_Pricer::price (src/Pricer/kwFd1d.cpp:68-157).  HBM traffic per PDE: 56 B of option
// parameters in, 8 B per priced option out.
//
// Partitioned Thomas ("SPIKE" with time-invariant spikes).  Thread k of P owns the M
// contiguous nodes k*M .. k*M+M-1 ("chunk k").  Because the LU factors never change, the two
// sweeps are first-order linear recurrences with constant multipliers, so for chunk k
//     y~_i = yl_i + Pp_i * Yin_k                  (yl: local forward sweep started from 0)
//     u~_i = ul_i + R_i * Yin_k + Q_i * Uin_k      (ul: local backward sweep of yl)
// where Pp (prefix products of a~), Q (suffix products of g~) and R (backward sweep of Pp)
// are precomputed, and Yin_k / Uin_k -- the true sweep values just outside the chunk -- come
// from two warp-level Kogge-Stone scans with precomputed multipliers (5 shuffle+FMA levels
// each) plus ONE __syncthreads per time step for the cross-warp carries:
//   * forward scan  S_k  = yl_end[k] + A_k S_{k-1}            (inclusive, lanes of one warp)
//   * backward scan on SHIFTED lanes: lane k holds chunk k+1's value,
//         T_k = (ul_0[k+1] + R0[k+1] S_k) + G_{k+1} T_{k+1},
//     so T_k is directly Uin_k and only the inclusive S_k is on the critical path;
//   * the cross-warp forward carry X_w enters the backward scan linearly through the
//     precomputed response H, so both warp-boundary values are exchanged at the same barrier
//     and every warp evaluates its two carries as dot products with precomputed rows.
// Per node-step: 2 local-sweep FMAs + 3 fix-up FMAs + 1 compare/select.
#pragma once
#include <type_traits>

#include "fd1d_common.cuh"
#include "tmem.cuh"

namespace kwfd1d {

constexpr unsigned FULL = 0xffffffffu;

template <int M, int P>
struct RegSmem {
    static constexpr int N = M * P;
    static constexpr int NW = P / 32;
    static constexpr int NWP = (NW + 1) & ~1;  // padded to even for 16-byte loads
    static constexpr int SCRATCH = (N > 8 * P) ? N : 8 * P;
    // per-warp exchange tables (doubles): rows cf, cb, ce [NW][NWP]; zf, zb [2][NWP]; AWf, AWb, H0 [NW]
    static constexpr int ZS = (NW + 2) * 32;   // exchange slots per parity buffer: [warp -1 .. NW][lane], pads stay 0
    static constexpr int WARP_TAB = 3 * NW * NWP + 2 * ZS + 3 * NW + 2 * P + 4;
    // doubles: xs[N] | scratch[SCRATCH] | proj[N] (optional) | dq[N] (optional) | tables
    static constexpr size_t bytes(bool proj_smem, bool dq_smem)
    {
        return sizeof(double) * (size_t)(N + SCRATCH + (proj_smem ? N : 0) + (dq_smem ? N : 0) + WARP_TAB);
    }
};

struct Mat2 {
    double m00, m01, m10, m11;
};

__device__ __forceinline__ Mat2 mat_mul(const Mat2& a, const Mat2& b)  // a * b
{
    Mat2 r;
    r.m00 = fma(a.m00, b.m00, a.m01 * b.m10);
    r.m01 = fma(a.m00, b.m01, a.m01 * b.m11);
    r.m10 = fma(a.m10, b.m00, a.m11 * b.m10);
    r.m11 = fma(a.m10, b.m01, a.m11 * b.m11);
    return r;
}

__device__ __forceinline__ void mat_normalise(Mat2& a)
{
    const double m = fmax(fmax(fabs(a.m00), fabs(a.m01)), fmax(fabs(a.m10), fabs(a.m11)));
    const double s = 1. / m;
    a.m00 *= s;
    a.m01 *= s;
    a.m10 *= s;
    a.m11 *= s;
}

__device__ __forceinline__ Mat2 mat_shfl_up(const Mat2& a, int d, int width = 32)
{
    Mat2 r;
    r.m00 = __shfl_up_sync(FULL, a.m00, d, width);
    r.m01 = __shfl_up_sync(FULL, a.m01, d, width);
    r.m10 = __shfl_up_sync(FULL, a.m10, d, width);
    r.m11 = __shfl_up_sync(FULL, a.m11, d, width);
    return r;
}

// Explicit shared-space accesses for the time loop: a 32-bit shared address kept in a register
// (generic pointers made ptxas re-derive the shared window base with S2UR inside the loop).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// march-type-generic shared accesses: the exchange slots keep their 8-byte stride for float too
template <class F>
struct Pair {
    F x, y;
};
template <class F>
__device__ __forceinline__ Pair<F> make_pair_t(F x, F y)
{
    Pair<F> p;
    p.x = x;
    p.y = y;
    return p;
}
template <class F>
__device__ __forceinline__ F lds_t(uint32_t a);
template <>
__device__ __forceinline__ double lds_t<double>(uint32_t a)
{
    return lds_f64(a);
}
template <>
__device__ __forceinline__ float lds_t<float>(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_t_impl(uint32_t a, double v) { sts_f64(a, v); }
__device__ __forceinline__ void sts_t_impl(uint32_t a, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
template <class F>
__device__ __forceinline__ void sts_t(uint32_t a, F v)
{
    sts_t_impl(a, v);
}
template <class F>
__device__ __forceinline__ Pair<F> lds_pair(uint32_t a)
{
    const double2 d = lds_v2f64(a);
    return make_pair_t<F>((F)d.x, (F)d.y);
}

// std::max(v, payoff) of the reference (src/Math/kwFd1d.cpp:131): (a < b) ? b : a.
// One DSETP + two selects; fmax() would add NaN canonicalisation the scheme does not need.
template <class F>
__device__ __forceinline__ F max_like_std(F a, F b)
{
    return (a < b) ? b : a;
}
// The same selection with the comparison on the integer pipe: for b >= +0 the order of the bit
// patterns as signed 64-bit integers is the order of the doubles (every negative a, and -0, sorts
// below +0; NaNs do not occur).  "Never project" is the floor -0.0 = INT64_MIN.  The only
// difference from the DSETP form: a = -0.0 against b = +0.0 yields +0.0, the same number.
__device__ __forceinline__ double max_like_icmp(double a, double b)
{
    return (__double_as_longlong(a) < __double_as_longlong(b)) ? b : a;
}
__device__ __forceinline__ float max_like_icmp(float a, float b)
{
    return (__float_as_int(a) < __float_as_int(b)) ? b : a;
}

// ---------------------------------------------------------------------------------------------
// Set-up shared by the CTA-per-PDE kernel below and the warp-per-PDE kernel (fd1d_warp.cuh): P
// threads, thread k owns the M nodes of chunk k.  Fills the x grid (xs, shared), the payoff v, the
// projection floor pj (no_floor where the reference does not project) and the pivot-scaled LU of
// B = 1 - dt/2 A:  a~ (a[0] = chunk-entry multiplier), g~ (g[M-1] = chunk-exit), D = 2/beta.
// `scr` is >= 8*P doubles of shared scratch.  Contains __syncthreads(): call from every thread of the CTA
// (groups of P threads work on different PDEs with their own xs / scr).
template <int M, int P>
__device__ __forceinline__ void setup_lu(const Fd1dBatch& B, const PdeScalars& sc, double no_floor, double* xs,
                                         double* scr, double (&v)[M], double (&pj)[M], double (&a)[M],
                                         double (&g)[M], double (&D)[M])
{
    constexpr int N = M * P;
    const int k = threadIdx.x % P;  // a CTA wider than P sets several PDEs up side by side (own xs / scr each)
    const int lane = k & 31;
    const int warp = k >> 5;
    const int xDim = B.xDim;
    double* s_bu_last = scr;           // [P]
    double* s_mat = scr + P;           // [4][P]
    double* s_bin = scr + 5 * P;       // [P] pivot just before each chunk
    double* s_ib_first = scr + 6 * P;  // [P]
    double* s_ib_last = scr + 7 * P;   // [P]
    // ---------------- set-up: grid, payoff --------------------------------------------
    double xl[M];  // this chunk's nodes stay in registers: only the two neighbours come back from shared memory
#pragma unroll
    for (int i = 0; i < M; ++i) {
        const int j = k * M + i;
        const double x = x_node(sc, B.density, j);
        xl[i] = x;
        xs[j] = x;
        double p = 0.;
        if (j < xDim) p = payoff_node(sc.put, x);
        v[i] = p;
        // projection skips the last node (src/Math/kwFd1d.cpp:130); European: never
        pj[i] = (sc.american && j < xDim - 1) ? p : no_floor;
    }
    __syncthreads();

    // ---------------- B rows, Moebius-composed pivots ---------------------------------
    {
        double bl[M], bb[M], bu[M];
        const double x_before = xs[k > 0 ? k * M - 1 : 0];
        const double x_after = xs[k < P - 1 ? k * M + M : N - 1];
        double inv_d = k > 0 ? 1. / (xl[0] - x_before) : 0.;  // 1 / (x_j - x_{j-1}), handed from row to row (b_row_chained)
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const int j = k * M + i;
            const double xm = i ? xl[i - 1] : x_before;           // = xs[j > 0 ? j - 1 : 0]
            const double xp = i < M - 1 ? xl[i + 1] : x_after;    // = xs[j < N - 1 ? j + 1 : N - 1]
            b_row_chained(sc, j, xDim, xm, xl[i], xp, inv_d, inv_d, bl[i], bb[i], bu[i]);
        }
        s_bu_last[k] = bu[M - 1];
        __syncthreads();
        const double bu_prev = k > 0 ? s_bu_last[k - 1] : 0.;

        // beta_j = b_j - c_j / beta_{j-1}, c_j = bl_j * bu_{j-1}: as a Moebius map on
        // (num; den) it is [[b_j, -c_j], [1, 0]]; compose the chunk's M maps
        {
            Mat2 m = {1., 0., 0., 1.};
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const double c = bl[i] * (i ? bu[i - 1] : bu_prev);
                Mat2 n;
                n.m00 = fma(bb[i], m.m00, -c * m.m10);
                n.m01 = fma(bb[i], m.m01, -c * m.m11);
                n.m10 = m.m00;
                n.m11 = m.m01;
                m = n;
                if ((i & 3) == 3) mat_normalise(m);
            }
            s_mat[k] = m.m00;
            s_mat[P + k] = m.m01;
            s_mat[2 * P + k] = m.m10;
            s_mat[3 * P + k] = m.m11;
        }
        __syncthreads();
        if (warp == 0) {
            constexpr int CPL = P / 32;  // chunk maps per lane
            Mat2 Lm = {1., 0., 0., 1.};
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int idx = lane * CPL + c;
                const Mat2 m = {s_mat[idx], s_mat[P + idx], s_mat[2 * P + idx], s_mat[3 * P + idx]};
                Lm = mat_mul(m, Lm);
                mat_normalise(Lm);
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Mat2 o = mat_shfl_up(Lm, d);
                if (lane >= d) {
                    Lm = mat_mul(Lm, o);
                    mat_normalise(Lm);
                }
            }
            const Mat2 E = mat_shfl_up(Lm, 1);
            double num = lane ? E.m00 : 1.;
            double den = lane ? E.m10 : 0.;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int idx = lane * CPL + c;
                s_bin[idx] = den == 0. ? CUDART_INF : num / den;
                const double nn = fma(s_mat[idx], num, s_mat[P + idx] * den);
                const double dd = fma(s_mat[2 * P + idx], num, s_mat[3 * P + idx] * den);
                const double s = 1. / fmax(fabs(nn), fabs(dd));
                num = nn * s;
                den = dd * s;
            }
        }
        __syncthreads();

        // The scan's pivots are a first guess only.  The recurrence contracts (d beta_j / d beta_{j-1} = c_j / beta_{j-1}^2
        // < 1), so run serially it forgets its rounding errors, while the composed maps accumulate theirs over the whole grid
        // (1e-13 relative at 4096 nodes: on stiff grids -- few time steps, dt/dx^2 in the thousands -- that alone moved prices
        // by 3e-9 against the reference).  Polish: every thread re-runs the reference's recurrence over its chunk from its
        // incoming pivot and hands the result to the next chunk, until no incoming pivot changes any more (chunk k is exact
        // after k sweeps at the latest; with the contraction a few sweeps do, ~25 on the stiffest grids).  The fixed point IS
        // the serial recurrence of src/Math/kwMath.cpp:30-38, bit for bit.
        double beta[M];  // the last sweep's pivots ARE the pivots (it ran from the final incoming pivot)
        {
            double pin = s_bin[k];
            __syncthreads();  // every thread holds its first guess before a neighbour overwrites it
#pragma unroll 1
            for (int sweep = 0; sweep < P + 2; ++sweep) {
                double prev = pin;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    // the reference's order (src/Math/kwMath.cpp:32-33): gam = au[j-1] / bet;  bet = a[j] - al[j] * gam
                    const double gam = (i ? bu[i - 1] : bu_prev) / prev;
                    prev = __dsub_rn(bb[i], __dmul_rn(bl[i], gam));
                    beta[i] = prev;
                }
                if (k < P - 1) s_bin[k + 1] = prev;
                __syncthreads();
                const double pnew = s_bin[k];
                const bool changed = __double_as_longlong(pnew) != __double_as_longlong(pin);
                pin = pnew;
                if (!__syncthreads_or(changed)) break;
            }
        }

        double ib[M];
#pragma unroll
        for (int i = 0; i < M; ++i) ib[i] = 1. / beta[i];
        s_ib_first[k] = ib[0];
        s_ib_last[k] = ib[M - 1];
        __syncthreads();
        const double ib_prev = k > 0 ? s_ib_last[k - 1] : 0.;
        const double ib_next = k < P - 1 ? s_ib_first[k + 1] : 0.;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const int j = k * M + i;
            a[i] = -bl[i] * (i ? ib[i - 1] : ib_prev);
            g[i] = -bu[i] * (i < M - 1 ? ib[i + 1] : ib_next);
            D[i] = j < xDim ? 2. * ib[i] : 0.;
        }
    }
}

// PROJ_SMEM / DQ_SMEM: keep the projection floor / the D*Q fix-up array in shared memory
// instead of registers (fewer registers -> more CTAs per SM, more shared-memory wavefronts).
// F: the arithmetic type of the time march.  double = the parity path (1e-9 absolute); float =
// FD1D.GPU.PRECISION f32: the whole set-up (grid, payoff, B rows, pivots, spikes, scan multipliers)
// stays in fp64 and only the march runs in fp32 (an all-fp32 set-up misses the 1e-4 bar).
// TMEM: the coefficient arrays g~, D, a~, D*R, D*Q and the payoff floor live in tensor memory
// (tmem.cuh) instead of registers and are streamed back with tcgen05.ld just before use: 128 instead
// of 168 registers per thread, i.e. 4 instead of 3 resident PDEs per SM, without touching the
// shared-memory pipe that the scans' shuffles already keep busy.
template <class F, int M, int P, int MINB, bool PROJ_SMEM, bool DQ_SMEM, bool ICMP = false, bool TMEM = false>
__global__ void __launch_bounds__(P, MINB) fd1d_reg_kernel(const Fd1dBatch B)
{
    static_assert(!TMEM || (std::is_same<F, double>::value && M == 8 && P == 128 && !PROJ_SMEM && !DQ_SMEM),
                  "TMEM variant: fp64, 8 nodes per thread, one warp per TMEM lane quarter");
    static_assert(M % 2 == 0 && P % 32 == 0, "M even, whole warps");
    static_assert(std::is_same<F, double>::value || (!PROJ_SMEM && !DQ_SMEM), "fp32 march keeps everything in registers");
    using L = RegSmem<M, P>;
    constexpr int N = L::N;
    constexpr int NW = L::NW;
    constexpr int NWP = L::NWP;
    constexpr int ZS = L::ZS;
    constexpr int M2 = M / 2;

    extern __shared__ double smem[];
    double* xs = smem;    // x grid, natural order
    double* scr = xs + N; // set-up exchange, then the final v
    double2* proj2 = reinterpret_cast<double2*>(scr + L::SCRATCH);  // [M2][P] pairs
    double2* dq2 = reinterpret_cast<double2*>(scr + L::SCRATCH + (PROJ_SMEM ? N : 0));
    double* tab = scr + L::SCRATCH + (PROJ_SMEM ? N : 0) + (DQ_SMEM ? N : 0);
    double* t_cf = tab;                  // [NW][NWP]  X_w  = sum_w' cf[w][w'] zf[w']
    double* t_cb = t_cf + NW * NWP;      // [NW][NWP]  Xb_w = sum_w' cb[w][w'] zb[w'] + ce[w][w'] zf[w']
    double* t_ce = t_cb + NW * NWP;      // [NW][NWP]
    double* zx = t_ce + NW * NWP;        // [2][NW + 2][32] exchange slots, double-buffered by step parity:
                                         // lane 31's slot = warp-end forward value, lane 0's = warp-start backward value
    double* AWf = zx + 2 * ZS;           // [NW] product of the forward chunk multipliers of warp w
    double* AWb = AWf + NW;              // [NW] same, backward
    double* H0 = AWb + NW;               // [NW] response of warp w's first backward value to X_w
    double* s_af4 = H0 + NW;             // [P] level-4 forward multipliers (general mode only)
    double* s_gb4 = s_af4 + P;           // [P] level-4 backward multipliers (general mode only)

    const int k = threadIdx.x;
    const int lane = k & 31;
    const int warp = k >> 5;
    const int xDim = B.xDim;
    const int nsteps = B.tDim - 1;

    // tensor memory: 128 columns per CTA = 64 doubles per thread (48 used), freed at kernel exit
    uint32_t tbase = 0;
    if constexpr (TMEM) {
        __shared__ uint32_t s_taddr;
        if (warp == 0) tmem::alloc128(smem_addr(&s_taddr));
        tmem::fence_before();
        __syncthreads();
        tmem::fence_after();
        tbase = s_taddr + ((uint32_t)(warp & 3) << 21);  // lane field (bits 31:16) = 32 * (warp % 4)
    }

    const uint32_t n_pde = batch_n_pde(B);
    for (uint32_t pde = blockIdx.x; pde < n_pde; pde += gridDim.x) {
        const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + pde) : pde;
        const kw_option opt = load_option(B.opts + rep);
        const PdeScalars sc = pde_scalars(opt, B);

        // ---------------- set-up: grid, payoff, B rows, Moebius-composed pivots -------------
        double v[M], pj[M];
        double a[M], g[M], D[M];  // a~ (a[0] = chunk-entry multiplier), g~ (g[M-1] = chunk-exit), 2/beta
        double DR[M], DQ[M];
        double Af[4], Gb[4], PWexf, PWbs, R0n, Hs, G0;
        setup_lu<M, P>(B, sc, ICMP ? -0. : -CUDART_INF, xs, scr, v, pj, a, g, D);
        if (PROJ_SMEM) {
#pragma unroll
            for (int c = 0; c < M2; ++c) proj2[c * P + k] = make_double2(pj[2 * c], pj[2 * c + 1]);
        }
        // spikes: Pp prefix products of a~, Q suffix products of g~, R = backward sweep of Pp
        bool bad_far = false, bad_l4 = false, bad_l3 = false, bad_l2 = false;  // "not negligible" votes
        {
            double Pp[M], Q[M], R[M];
            Pp[0] = a[0];
#pragma unroll
            for (int i = 1; i < M; ++i) Pp[i] = a[i] * Pp[i - 1];
            Q[M - 1] = g[M - 1];
#pragma unroll
            for (int i = M - 2; i >= 0; --i) Q[i] = g[i] * Q[i + 1];
            R[M - 1] = Pp[M - 1];
#pragma unroll
            for (int i = M - 2; i >= 0; --i) R[i] = fma(g[i], R[i + 1], Pp[i]);
#pragma unroll
            for (int i = 0; i < M; ++i) {
                DR[i] = D[i] * R[i];
                DQ[i] = D[i] * Q[i];
            }
            // ---- forward Kogge-Stone multipliers (chunk products A_k = Pp[M-1])
            double A = Pp[M - 1];
            double af4 = 0., gb4 = 0.;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int s = 1 << d;
                const double o = __shfl_up_sync(FULL, A, s);
                const double m = lane >= s ? A : 0.;
                if (d < 4)
                    Af[d] = m;
                else
                    af4 = m;
                if (lane >= s) A *= o;
            }
            const double PWf = A;  // product warp-start .. k (inclusive)
            {
                const double ex = __shfl_up_sync(FULL, A, 1);
                PWexf = lane ? ex : 1.;
            }
            if (lane == 31) AWf[warp] = A;
            // ---- backward scan on shifted lanes: lane k carries chunk k+1
            G0 = Q[0];  // this chunk's backward multiplier
            const double R0 = R[0];
            {
                const double gn = __shfl_down_sync(FULL, G0, 1);
                const double rn = __shfl_down_sync(FULL, R0, 1);
                R0n = lane < 31 ? rn : 0.;
                double G = lane < 31 ? gn : 0.;  // Gn_k = G_{k+1}
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int s = 1 << d;
                    const double o = __shfl_down_sync(FULL, G, s);
                    const double m = lane < 32 - s ? G : 0.;
                    if (d < 4)
                        Gb[d] = m;
                    else
                        gb4 = m;
                    if (lane < 32 - s) G *= o;
                }
                // G = prod_{i=k}^{31} Gn_i, which is 0 through Gn_31 = 0; the carry multiplier
                // PWbs_k = prod_{i=k}^{30} Gn_i needs the product without that last factor
                double Gx = lane < 31 ? gn : 1.;
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int s = 1 << d;
                    const double o = __shfl_down_sync(FULL, Gx, s);
                    if (lane < 32 - s) Gx *= o;
                }
                PWbs = Gx;  // lane 31: 1
                // Hs: shifted backward scan of the response R0n * PWf to a unit forward carry X_w
                double H = R0n * PWf;
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const double o = __shfl_down_sync(FULL, H, 1 << d);
                    H = fma(d < 4 ? Gb[d] : gb4, o, H);
                }
                Hs = H;
                if (lane == 0) {
                    AWb[warp] = G0 * PWbs;       // prod_{j=0}^{31} G_j
                    H0[warp] = fma(G0, Hs, R0);  // d(T_0)/d(X_w)
                }
            }
            s_af4[k] = af4;
            s_gb4[k] = gb4;
            if (DQ_SMEM) {
#pragma unroll
                for (int c = 0; c < M2; ++c) dq2[c * P + k] = make_double2(DQ[2 * c], DQ[2 * c + 1]);
            }
            // ---- negligibility votes (DESIGN.md "Truncation").  A carry term may be dropped when its
            // absolute size, relative to the local solution scale max(1, e^x), accumulated over all
            // time steps, stays below 2^-56: |multiplier| * Bmax * growth * tDim <= 2^-56, where
            // Bmax bounds the pivot scaling of the carried values and growth = 1 for puts (|v| <= 1)
            // and e^(x_source - x_here) for calls (v <= e^x) in the backward direction.
            double bmax = 0.;
#pragma unroll
            for (int i = 0; i < M; ++i) bmax = fmax(bmax, D[i] != 0. ? fabs(2. / D[i]) : 1.);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) bmax = fmax(bmax, __shfl_xor_sync(FULL, bmax, d));
            if (lane == 0) scr[warp] = bmax;  // scr is free again (all set-up exchanges are done)
            __syncthreads();
            bmax = scr[0];
#pragma unroll
            for (int w = 1; w < NW; ++w) bmax = fmax(bmax, scr[w]);
            // fp32 march: anything below 2^-30 of the local scale over the whole march is far inside its
            // own rounding (2^-24 per operation)
            const double tol = (std::is_same<F, double>::value ? 0x1p-56 : 0x1p-30) / (bmax * (double)B.tDim);
            const int j_src = min((warp + 1) * 32 * M + M - 1, xDim - 1);
            const double growth =
                sc.put ? 1. : exp(fmax(0., xs[j_src]) - fmax(0., xs[min(k * M, xDim - 1)]));
            bad_l4 = !(fabs(af4) <= tol) || !(fabs(gb4) * growth <= tol);
            bad_l3 = !(fabs(Af[3]) <= tol) || !(fabs(Gb[3]) * growth <= tol);
            bad_l2 = !(fabs(Af[2]) <= tol) || !(fabs(Gb[2]) * growth <= tol);
            if (lane == 31) bad_far = !(fabs(A) <= tol);
            if (lane == 0) {
                const double gfar = sc.put ? 1. : exp(fmax(0., xs[xDim - 1]) - fmax(0., xs[min(k * M, xDim - 1)]));
                bad_far = !(fabs(G0 * PWbs) * gfar <= tol);
            }
        }
        const int any_far = __syncthreads_or(bad_far);
        const int any_l4 = __syncthreads_or(bad_l4);
        const int any_l3 = __syncthreads_or(bad_l3);
        const int any_l2 = __syncthreads_or(bad_l2);
        // carry modes.  0: every carry term; 1: only nearest-warp carries, all 5 scan levels;
        // 2 / 3 / 4: nearest-warp carries and 4 / 3 / 2 scan levels (higher levels proven negligible).
        // B.max_mode caps it (FD1D.GPU.EXACT = 2 -> mode 0 everywhere, 1 -> at most mode 1).
        int mode = 0;
        if (!(NW > 1 && any_far)) mode = any_l4 ? 1 : (any_l3 ? 2 : (any_l2 ? 3 : 4));
        if (mode > B.max_mode) mode = B.max_mode;
        if (NW == 1 && mode == 0) mode = 1;  // no cross-warp carries to be exact about
        // cross-warp rows (mode 0 only): thread w < NW fills row w
        if (NW > 1 && mode == 0) {
            if (k < NW) {
                const int w = k;
                // cf[w][w'] = prod_{w' < w'' < w} AWf[w'']  (w' < w), else 0
                double c = 1.;
                for (int wp = NWP - 1; wp >= 0; --wp) {
                    double val = 0.;
                    if (wp < w) {
                        val = c;
                        c *= AWf[wp];
                    }
                    t_cf[w * NWP + wp] = val;
                }
                // cb[w][w'] = prod_{w < w'' < w'} AWb[w'']  (w' > w), else 0
                c = 1.;
                for (int wp = 0; wp < NWP; ++wp) {
                    double val = 0.;
                    if (wp > w && wp < NW) {
                        val = c;
                        c *= AWb[wp];
                    }
                    t_cb[w * NWP + wp] = val;
                }
            }
            __syncthreads();
            if (k < NW) {
                const int w = k;
                // ce[w][w''] = sum_{w' > max(w, w'')} cb[w][w'] * H0[w'] * cf[w'][w'']
                for (int ws = 0; ws < NWP; ++ws) {
                    double acc = 0.;
                    for (int wp = (w > ws ? w : ws) + 1; wp < NW; ++wp)
                        acc = fma(t_cb[w * NWP + wp] * H0[wp], t_cf[wp * NWP + ws], acc);
                    t_ce[w * NWP + ws] = acc;
                }
            }
        }
        // exchange slots, zero pads for the edge warps (slot rows 0 and NW + 1 of both buffers)
        for (int i = k; i < 2 * ZS; i += P) zx[i] = 0.;
        __syncthreads();
        const double H0n = (NW > 1 && warp < NW - 1) ? H0[warp + 1] : 0.;
        const uint32_t a_proj = smem_addr(proj2 + k);
        const uint32_t a_dq = smem_addr(dq2 + k);
        const uint32_t a_mine = smem_addr(zx + (warp + 1) * 32 + lane);  // this lane's slot, parity 0
        const uint32_t a_zx = smem_addr(zx);
        const uint32_t a_rows = smem_addr(t_cf + warp * NWP);

        // march-typed copies (identity for the fp64 march)
        F vm[M], am[M], gm[M], Dm[M], DRm[M], DQm[M], pm[M], Afm[4], Gbm[4];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            vm[i] = (F)v[i];
            am[i] = (F)a[i];
            gm[i] = (F)g[i];
            Dm[i] = (F)D[i];
            DRm[i] = (F)DR[i];
            DQm[i] = (F)DQ[i];
            pm[i] = (F)pj[i];
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            Afm[d] = (F)Af[d];
            Gbm[d] = (F)Gb[d];
        }
        const F PWexfm = (F)PWexf, PWbsm = (F)PWbs, R0nm = (F)R0n, Hsm = (F)Hs, G0m = (F)G0, H0nm = (F)H0n;

        // ---------------- time march: tDim-1 steps, one barrier each ----------------------
        auto march = [&](auto mode_c) {
            constexpr int MODE = decltype(mode_c)::value;
            constexpr int NLEV = MODE <= 1 ? 5 : 6 - MODE;  // scan levels kept
            F af4 = F(0), gb4 = F(0);
            if (NLEV == 5) {
                af4 = (F)s_af4[k];
                gb4 = (F)s_gb4[k];
            }
            // local forward sweep from 0 for the first step (later ones are fused into the loop tail)
            F y[M];
            y[0] = vm[0];
#pragma unroll
            for (int i = 1; i < M; ++i) y[i] = fma(am[i], y[i - 1], vm[i]);
            // software pipeline: the level-0 shuffle of the forward scan is issued at the END of the
            // previous iteration, so the local backward sweep runs in its shadow
            F o0 = __shfl_up_sync(FULL, y[M - 1], 1);
            for (int step = 0; step < nsteps; ++step) {
                // inclusive forward scan of the chunk-end values inside the warp
                F S = fma(Afm[0], o0, y[M - 1]);
#pragma unroll
                for (int d = 1; d < 4; ++d) {
                    if (d < NLEV) {
                        const F o = __shfl_up_sync(FULL, S, 1 << d);
                        S = fma(Afm[d], o, S);
                    }
                }
                if (NLEV == 5) {
                    const F o = __shfl_up_sync(FULL, S, 16);
                    S = fma(af4, o, S);
                }
                // local backward sweep of the local forward result; r_i = D_i ul_i - v_i on the fly
#pragma unroll
                for (int i = M - 2; i >= 0; --i) y[i] = fma(gm[i], y[i + 1], y[i]);
#pragma unroll
                for (int i = 0; i < M; ++i) vm[i] = fma(Dm[i], y[i], -vm[i]);
                // shifted backward scan: lane k holds chunk k+1
                const F uln = __shfl_down_sync(FULL, y[0], 1);
                F T = fma(R0nm, S, lane < 31 ? uln : F(0));
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    if (d < NLEV) {
                        const F o = __shfl_down_sync(FULL, T, 1 << d);
                        T = fma(Gbm[d], o, T);
                    }
                }
                if (NLEV == 5) {
                    const F o = __shfl_down_sync(FULL, T, 16);
                    T = fma(gb4, o, T);
                }
                F Sm1 = __shfl_up_sync(FULL, S, 1);
                if (lane == 0) Sm1 = F(0);

                F Xw = F(0), Xbw = F(0);
                if (NW > 1) {
                    // every lane stores (no divergent branch); only lanes 0 and 31 are read back
                    const uint32_t par = (uint32_t)(step & 1) * (ZS * 8);
                    sts_t<F>(a_mine + par, lane == 0 ? fma(G0m, T, y[0]) : S);
                    __syncthreads();
                    const uint32_t zrow = a_zx + par + (uint32_t)warp * 256;  // slot row of warp - 1
                    if (MODE == 0) {
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            const F f = lds_t<F>(a_zx + par + (w + 1) * 256 + 31 * 8);
                            const F b = lds_t<F>(a_zx + par + (w + 1) * 256);
                            const F cf = (F)lds_f64(a_rows + w * 8);
                            const F cb = (F)lds_f64(a_rows + (NW * NWP + w) * 8);
                            const F ce = (F)lds_f64(a_rows + (2 * NW * NWP + w) * 8);
                            Xw = fma(cf, f, Xw);
                            Xbw = fma(cb, b, Xbw);
                            Xbw = fma(ce, f, Xbw);
                        }
                    } else {
                        // nearest-warp carries only: X_w = zf[w-1]; Xb_w = zb[w+1] + H0[w+1] * zf[w]
                        const F fm1 = lds_t<F>(zrow + 31 * 8);
                        const F fw = lds_t<F>(zrow + 256 + 31 * 8);
                        const F bp1 = lds_t<F>(zrow + 512);
                        Xw = fm1;
                        Xbw = fma(H0nm, fw, bp1);
                    }
                }
                const F Yin = fma(PWexfm, Xw, Sm1);
                const F Uin = fma(PWbsm, Xbw, fma(Hsm, Xw, T));
                // fix-up + projection, fused with the NEXT step's local forward sweep so that the
                // sweep's dependent chain hides behind the independent per-node fix-ups
#pragma unroll
                for (int c = 0; c < M2; ++c) {
                    Pair<F> pp, qq;
                    if (PROJ_SMEM)
                        pp = lds_pair<F>(a_proj + c * (P * 16));
                    else
                        pp = make_pair_t<F>(pm[2 * c], pm[2 * c + 1]);
                    if (DQ_SMEM)
                        qq = lds_pair<F>(a_dq + c * (P * 16));
                    else
                        qq = make_pair_t<F>(DQm[2 * c], DQm[2 * c + 1]);
                    {
                        const int i = 2 * c;
                        F r = fma(DRm[i], Yin, vm[i]);
                        r = fma(qq.x, Uin, r);
                        vm[i] = ICMP ? max_like_icmp(r, pp.x) : max_like_std(r, pp.x);
                        y[i] = i ? fma(am[i], y[i - 1], vm[i]) : vm[i];
                    }
                    {
                        const int i = 2 * c + 1;
                        F r = fma(DRm[i], Yin, vm[i]);
                        r = fma(qq.y, Uin, r);
                        vm[i] = ICMP ? max_like_icmp(r, pp.y) : max_like_std(r, pp.y);
                        y[i] = fma(am[i], y[i - 1], vm[i]);
                    }
                }
                o0 = __shfl_up_sync(FULL, y[M - 1], 1);
            }
        };
        // ---------------- TMEM variant of the march ----------------------------------------
        // columns (2 per double): [0,16) g~[0..7] | [16,32) D[0..7] | [32+16c, 48+16c) node pair c:
        // a~[2c], a~[2c+1], DR[2c], DR[2c+1], DQ[2c], DQ[2c+1], p[2c], p[2c+1]
        if constexpr (TMEM) {
            double t8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = g[i];
            tmem::st8(tbase, t8);
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = D[i];
            tmem::st8(tbase + 16, t8);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                t8[0] = a[2 * c];
                t8[1] = a[2 * c + 1];
                t8[2] = DR[2 * c];
                t8[3] = DR[2 * c + 1];
                t8[4] = DQ[2 * c];
                t8[5] = DQ[2 * c + 1];
                t8[6] = pj[2 * c];
                t8[7] = pj[2 * c + 1];
                tmem::st8(tbase + 32 + 16 * c, t8);
            }
            tmem::wait_st();
        }
        auto march_tmem = [&](auto mode_c) {
            constexpr int MODE = decltype(mode_c)::value;
            constexpr int NLEV = MODE <= 1 ? 5 : 6 - MODE;
            double af4 = 0., gb4 = 0.;
            if (NLEV == 5) {
                af4 = s_af4[k];
                gb4 = s_gb4[k];
            }
            double w[M], y[M];
#pragma unroll
            for (int i = 0; i < M; ++i) w[i] = v[i];
            // first local forward sweep: the a~ come from the pair blocks
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double q[8];
                tmem::ld8(tbase + 32 + 16 * c, q);
                tmem::wait_ld_dep(q);
                y[2 * c] = c ? fma(q[0], y[2 * c - 1], w[2 * c]) : w[0];
                y[2 * c + 1] = fma(q[1], y[2 * c], w[2 * c + 1]);
            }
            double o0 = __shfl_up_sync(FULL, y[M - 1], 1);
            double g8[8];
            tmem::ld8(tbase, g8);
            for (int step = 0; step < nsteps; ++step) {
                double S = fma(Af[0], o0, y[M - 1]);
#pragma unroll
                for (int d = 1; d < 4; ++d) {
                    if (d < NLEV) {
                        const double o = __shfl_up_sync(FULL, S, 1 << d);
                        S = fma(Af[d], o, S);
                    }
                }
                if (NLEV == 5) {
                    const double o = __shfl_up_sync(FULL, S, 16);
                    S = fma(af4, o, S);
                }
                tmem::wait_ld_dep(g8);
                double d8[8];
                tmem::ld8(tbase + 16, d8);
#pragma unroll
                for (int i = M - 2; i >= 0; --i) y[i] = fma(g8[i], y[i + 1], y[i]);
                tmem::wait_ld_dep(d8);
#pragma unroll
                for (int i = 0; i < M; ++i) w[i] = fma(d8[i], y[i], -w[i]);
                double q[4][8];
                tmem::ld8(tbase + 32, q[0]);
                const double uln = __shfl_down_sync(FULL, y[0], 1);
                double T = fma(R0n, S, lane < 31 ? uln : 0.);
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    if (d < NLEV) {
                        const double o = __shfl_down_sync(FULL, T, 1 << d);
                        T = fma(Gb[d], o, T);
                    }
                }
                if (NLEV == 5) {
                    const double o = __shfl_down_sync(FULL, T, 16);
                    T = fma(gb4, o, T);
                }
                double Sm1 = __shfl_up_sync(FULL, S, 1);
                if (lane == 0) Sm1 = 0.;
                double Xw = 0., Xbw = 0.;
                {
                    const uint32_t par = (uint32_t)(step & 1) * (ZS * 8);
                    sts_f64(a_mine + par, lane == 0 ? fma(G0, T, y[0]) : S);
                    __syncthreads();
                    const uint32_t zrow = a_zx + par + (uint32_t)warp * 256;
                    if (MODE == 0) {
#pragma unroll
                        for (int ww = 0; ww < NW; ++ww) {
                            const double f = lds_f64(a_zx + par + (ww + 1) * 256 + 31 * 8);
                            const double b = lds_f64(a_zx + par + (ww + 1) * 256);
                            const double cf = lds_f64(a_rows + ww * 8);
                            const double cb = lds_f64(a_rows + (NW * NWP + ww) * 8);
                            const double ce = lds_f64(a_rows + (2 * NW * NWP + ww) * 8);
                            Xw = fma(cf, f, Xw);
                            Xbw = fma(cb, b, Xbw);
                            Xbw = fma(ce, f, Xbw);
                        }
                    } else {
                        const double fm1 = lds_f64(zrow + 31 * 8);
                        const double fw = lds_f64(zrow + 256 + 31 * 8);
                        const double bp1 = lds_f64(zrow + 512);
                        Xw = fm1;
                        Xbw = fma(H0n, fw, bp1);
                    }
                }
                const double Yin = fma(PWexf, Xw, Sm1);
                const double Uin = fma(PWbs, Xbw, fma(Hs, Xw, T));
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tmem::wait_ld_dep(q[c]);
                    if (c < 3)
                        tmem::ld8(tbase + 32 + 16 * (c + 1), q[c + 1]);
                    else
                        tmem::ld8(tbase, g8);  // next step's g~
                    {
                        const int i = 2 * c;
                        double r = fma(q[c][2], Yin, w[i]);
                        r = fma(q[c][4], Uin, r);
                        w[i] = ICMP ? max_like_icmp(r, q[c][6]) : max_like_std(r, q[c][6]);
                        y[i] = i ? fma(q[c][0], y[i - 1], w[i]) : w[i];
                    }
                    {
                        const int i = 2 * c + 1;
                        double r = fma(q[c][3], Yin, w[i]);
                        r = fma(q[c][5], Uin, r);
                        w[i] = ICMP ? max_like_icmp(r, q[c][7]) : max_like_std(r, q[c][7]);
                        y[i] = fma(q[c][1], y[i - 1], w[i]);
                    }
                }
                o0 = __shfl_up_sync(FULL, y[M - 1], 1);
            }
            tmem::wait_ld_dep(g8);  // nothing in flight when the next PDE overwrites the arrays
#pragma unroll
            for (int i = 0; i < M; ++i) vm[i] = w[i];
        };
        if constexpr (TMEM) {
            switch (mode) {
                case 0: march_tmem(std::integral_constant<int, 0>{}); break;
                case 1: march_tmem(std::integral_constant<int, 1>{}); break;
                case 2: march_tmem(std::integral_constant<int, 2>{}); break;
                case 3: march_tmem(std::integral_constant<int, 3>{}); break;
                default: march_tmem(std::integral_constant<int, 4>{}); break;
            }
        } else {
            switch (mode) {
                case 0: march(std::integral_constant<int, 0>{}); break;
                case 1: march(std::integral_constant<int, 1>{}); break;
                case 2: march(std::integral_constant<int, 2>{}); break;
                case 3: march(std::integral_constant<int, 3>{}); break;
                default: march(std::integral_constant<int, 4>{}); break;
            }
        }
        if (k == 0) atomicAdd(&B.status[2 + mode], 1u);

        // ---------------- epilogue: interpolate every option of this chain -----------------
        __syncthreads();
#pragma unroll
        for (int i = 0; i < M; ++i) scr[k * M + i] = (double)vm[i];
        __syncthreads();
        {
            uint32_t q0, q1;
            chain_range(B, pde, q0, q1);
            for (uint32_t q = q0 + k; q < q1; q += P) {
                const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                price_option(B, oi, [&](int j) { return xs[j]; }, [&](int j) { return scr[j]; });
            }
        }
        __syncthreads();
    }
    if constexpr (TMEM) {
        tmem::fence_before();
        __syncthreads();
        if (warp == 0) tmem::dealloc128(tbase & 0xffffu);
    }
}

}  // namespace kwfd1d
