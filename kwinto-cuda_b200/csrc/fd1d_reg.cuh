// fd1d_reg.cuh -- Layout B: one CTA per PDE, the whole time march in one launch, x-grid
// coefficients in REGISTERS (payoff, and optionally one fix-up array, in shared memory).
//
// Replaces Fd1d::solve/solveOne + solveTridiagonal + Fd1d::value of the reference
// (src/Math/kwFd1d.cpp:11-158, src/Math/kwMath.cpp:16-49) and the grid/payoff set-up of
// Fd1d_Pricer::price (src/Pricer/kwFd1d.cpp:68-157).  HBM traffic per PDE: 56 B of option
// parameters in, 8 B per priced option out.
//
// Partitioned Thomas ("SPIKE" with time-invariant spikes).  Thread k of P owns the M
// contiguous nodes k*M .. k*M+M-1.  Because the LU factors never change, the two sweeps are
// first-order linear recurrences with constant multipliers, so for thread k
//     y~_i = yl_i + Pp_i * Yin_k                  (yl: local forward sweep from 0)
//     u~_i = ul_i + R_i * Yin_k + Q_i * Uin_k      (ul: local backward sweep of yl)
// where Pp (prefix products of a~), Q (suffix products of g~) and R (backward sweep of Pp)
// are precomputed, and Yin_k / Uin_k -- the true sweep values just outside the chunk -- come
// from two warp-level Kogge-Stone scans with precomputed multipliers (5 shuffle+FMA levels
// each) plus ONE __syncthreads per time step for the cross-warp carries: the backward scan is
// started with the warp-local forward carry and corrected afterwards with the precomputed
// response H of the backward scan to the cross-warp forward carry.
// Per node-step: 2 local-sweep FMAs + 3 fix-up FMAs + 1 max; per thread-step ~25 scan FMAs.
#pragma once
#include "fd1d_common.cuh"

namespace kwfd1d {

constexpr unsigned FULL = 0xffffffffu;

template <int M, int P>
struct RegSmem {
    static constexpr int N = M * P;
    static constexpr int NW = P / 32;
    static constexpr int SCRATCH = (N > 8 * P) ? N : 8 * P;
    // doubles: xs[N] | scratch[SCRATCH] | proj[N] | dq[N] (optional) | per-warp tables
    static constexpr int WARP_TAB = 3 * NW + 4 * NW;  // AWf, AWb, H0, zf[2][NW], zb[2][NW]
    static constexpr size_t bytes(bool dq_smem)
    {
        return sizeof(double) * (size_t)(N + SCRATCH + N + (dq_smem ? N : 0) + WARP_TAB + 8);
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

__device__ __forceinline__ Mat2 mat_shfl_up(const Mat2& a, int d)
{
    Mat2 r;
    r.m00 = __shfl_up_sync(FULL, a.m00, d);
    r.m01 = __shfl_up_sync(FULL, a.m01, d);
    r.m10 = __shfl_up_sync(FULL, a.m10, d);
    r.m11 = __shfl_up_sync(FULL, a.m11, d);
    return r;
}

// DQ_SMEM: keep the D*Q fix-up array in shared memory instead of registers (lower register
// count -> one more CTA per SM, at the price of shared-memory bandwidth).
template <int M, int P, int MINB, bool DQ_SMEM>
__global__ void __launch_bounds__(P, MINB) fd1d_reg_kernel(const Fd1dBatch B)
{
    static_assert(M % 2 == 0 && P % 32 == 0, "M even, whole warps");
    using L = RegSmem<M, P>;
    constexpr int N = L::N;
    constexpr int NW = L::NW;
    constexpr int M2 = M / 2;

    extern __shared__ double smem[];
    double* xs = smem;                                   // x grid, natural order
    double* scr = xs + N;                                // set-up exchange, then final v
    double2* proj2 = reinterpret_cast<double2*>(scr + L::SCRATCH);  // [M2][P] pairs
    double2* dq2 = reinterpret_cast<double2*>(scr + L::SCRATCH + N);
    double* tab = scr + L::SCRATCH + N + (DQ_SMEM ? N : 0);
    double* AWf = tab;            // product of forward chunk multipliers over warp w
    double* AWb = tab + NW;       // same, backward
    double* H0 = tab + 2 * NW;    // response of warp w's first backward value to its forward carry
    double* zf = tab + 3 * NW;    // [2][NW] warp-end forward values, double-buffered by step parity
    double* zb = tab + 5 * NW;    // [2][NW] warp-start backward values

    // scratch sub-arrays used during set-up
    double* s_bu_last = scr;          // [P]
    double* s_mat = scr + P;          // [4][P]
    double* s_bin = scr + 5 * P;      // [P] pivot just before each chunk
    double* s_ib_first = scr + 6 * P; // [P]
    double* s_ib_last = scr + 7 * P;  // [P]

    const int k = threadIdx.x;
    const int lane = k & 31;
    const int warp = k >> 5;
    const int xDim = B.xDim;
    const int nsteps = B.tDim - 1;

    for (uint32_t pde = blockIdx.x; pde < B.n_pde; pde += gridDim.x) {
        const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + pde) : pde;
        const kw_option opt = load_option(B.opts + rep);
        const PdeScalars sc = pde_scalars(opt, B);

        // ---------------- set-up: grid, payoff --------------------------------------------
        double v[M];
        {
            double pj[M];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const int j = k * M + i;
                const double x = x_node(sc, B.density, j);
                xs[j] = x;
                double p = 0.;
                if (j < xDim) p = payoff_node(sc.put, x);
                v[i] = p;
                // projection skips the last node (src/Math/kwFd1d.cpp:130); European: never
                pj[i] = (sc.american && j < xDim - 1) ? p : -CUDART_INF;
            }
#pragma unroll
            for (int c = 0; c < M2; ++c) proj2[c * P + k] = make_double2(pj[2 * c], pj[2 * c + 1]);
        }
        __syncthreads();

        // ---------------- B rows, Moebius-composed pivots ---------------------------------
        double a[M], g[M], D[M];  // a~ (a[0] = chunk-entry multiplier), g~ (g[M-1] = chunk-exit), 2/beta
        double DR[M], DQ[M];
        double Af[5], Gb[5], PWexf, PWexb, R0, Hp1;
        {
            double bl[M], bb[M], bu[M];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const int j = k * M + i;
                const double xm = xs[j > 0 ? j - 1 : 0];
                const double xp = xs[j < N - 1 ? j + 1 : N - 1];
                b_row(sc, j, xDim, xm, xs[j], xp, bl[i], bb[i], bu[i]);
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

            // pivots inside the chunk, in the reference's order (src/Math/kwMath.cpp:32-33):
            // gam = au[j-1] / bet;  bet = a[j] - al[j] * gam
            double ib[M];
            {
                double prev = s_bin[k];
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const double gam = (i ? bu[i - 1] : bu_prev) / prev;
                    const double beta = __dsub_rn(bb[i], __dmul_rn(bl[i], gam));
                    ib[i] = 1. / beta;
                    prev = beta;
                }
            }
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
        // spikes: Pp prefix products of a~, Q suffix products of g~, R = backward sweep of Pp
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
            R0 = R[0];
            // Kogge-Stone multipliers inside the warp (forward: chunk products A = Pp[M-1])
            double A = Pp[M - 1];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int s = 1 << d;
                const double o = __shfl_up_sync(FULL, A, s);
                Af[d] = lane >= s ? A : 0.;
                if (lane >= s) A *= o;
            }
            {
                const double ex = __shfl_up_sync(FULL, A, 1);
                PWexf = lane ? ex : 1.;
            }
            if (lane == 31) AWf[warp] = A;
            double G = Q[0];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int s = 1 << d;
                const double o = __shfl_down_sync(FULL, G, s);
                Gb[d] = lane < 32 - s ? G : 0.;
                if (lane < 32 - s) G *= o;
            }
            {
                const double ex = __shfl_down_sync(FULL, G, 1);
                PWexb = lane < 31 ? ex : 1.;
            }
            if (lane == 0) AWb[warp] = G;
            // H: backward in-warp scan of the response R0 * PWexf to a unit cross-warp forward carry
            double H = R0 * PWexf;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const double o = __shfl_down_sync(FULL, H, 1 << d);
                H = fma(Gb[d], o, H);
            }
            {
                const double nx = __shfl_down_sync(FULL, H, 1);
                Hp1 = lane < 31 ? nx : 0.;
            }
            if (lane == 0) H0[warp] = H;
            if (DQ_SMEM) {
#pragma unroll
                for (int c = 0; c < M2; ++c) dq2[c * P + k] = make_double2(DQ[2 * c], DQ[2 * c + 1]);
            }
        }
        __syncthreads();

        // ---------------- time march: tDim-1 steps, one barrier each ----------------------
        for (int step = 0; step < nsteps; ++step) {
            double y[M];
            y[0] = v[0];
#pragma unroll
            for (int i = 1; i < M; ++i) y[i] = fma(a[i], y[i - 1], v[i]);
            double S = y[M - 1];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const double o = __shfl_up_sync(FULL, S, 1 << d);
                S = fma(Af[d], o, S);
            }
            double Sm1 = __shfl_up_sync(FULL, S, 1);
            if (lane == 0) Sm1 = 0.;
            // local backward sweep of the local forward result
#pragma unroll
            for (int i = M - 2; i >= 0; --i) y[i] = fma(g[i], y[i + 1], y[i]);
            double T = fma(R0, Sm1, y[0]);
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const double o = __shfl_down_sync(FULL, T, 1 << d);
                T = fma(Gb[d], o, T);
            }
            double Tp1 = __shfl_down_sync(FULL, T, 1);
            if (lane == 31) Tp1 = 0.;

            double Xw = 0., Xbw = 0.;
            if (NW > 1) {
                const int par = (step & 1) * NW;
                if (lane == 31) zf[par + warp] = S;
                if (lane == 0) zb[par + warp] = T;
                __syncthreads();
                // forward carries into every warp, then the corrected backward warp-start
                // values, then the backward carry into this warp
                double Xs[NW];
                double X = 0.;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    Xs[w] = X;
                    if (w == warp) Xw = X;
                    X = fma(AWf[w], X, zf[par + w]);
                }
                double Xb = 0.;
#pragma unroll
                for (int w = NW - 1; w >= 0; --w) {
                    if (w == warp) Xbw = Xb;
                    const double zt = fma(Xs[w], H0[w], zb[par + w]);
                    Xb = fma(AWb[w], Xb, zt);
                }
            }
            const double Yin = fma(PWexf, Xw, Sm1);
            const double Uin = fma(PWexb, Xbw, fma(Hp1, Xw, Tp1));
#pragma unroll
            for (int c = 0; c < M2; ++c) {
                const double2 pp = proj2[c * P + k];
                double2 qq;
                if (DQ_SMEM)
                    qq = dq2[c * P + k];
                else
                    qq = make_double2(DQ[2 * c], DQ[2 * c + 1]);
                {
                    const int i = 2 * c;
                    double r = fma(D[i], y[i], -v[i]);
                    r = fma(DR[i], Yin, r);
                    r = fma(qq.x, Uin, r);
                    v[i] = fmax(r, pp.x);
                }
                {
                    const int i = 2 * c + 1;
                    double r = fma(D[i], y[i], -v[i]);
                    r = fma(DR[i], Yin, r);
                    r = fma(qq.y, Uin, r);
                    v[i] = fmax(r, pp.y);
                }
            }
        }

        // ---------------- epilogue: interpolate every option of this chain -----------------
        __syncthreads();
#pragma unroll
        for (int i = 0; i < M; ++i) scr[k * M + i] = v[i];
        __syncthreads();
        {
            const uint32_t q0 = B.csr_start ? __ldg(B.csr_start + pde) : pde;
            const uint32_t q1 = B.csr_start ? __ldg(B.csr_start + pde + 1) : pde + 1;
            for (uint32_t q = q0 + k; q < q1; q += P) {
                const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                price_option(B, oi, [&](int j) { return xs[j]; }, [&](int j) { return scr[j]; });
            }
        }
        __syncthreads();
    }
}

}  // namespace kwfd1d
