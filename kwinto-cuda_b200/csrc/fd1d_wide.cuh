// fd1d_wide.cuh -- Layout W for grids wider than one warp can hold: 1024 < xDim <= 4096.
//
// The march is fd1d_warp.cuh's (every lane owns 4 chunks of 8 nodes, a~, g~, D, p in tensor memory, v
// in registers, true sweeps from the chunk-boundary values, chunk pairs in lock step); a PDE now
// spans NWP = 2 or 4 warps, i.e. half a CTA or a whole one.  What is new:
//   * the cross-warp carries.  Inside a warp the Kogge-Stone scans run as if nothing flowed in; the
//     true values differ by (product of the lane multipliers since the warp boundary) x (true value
//     at the boundary), and the boundary values obey the same first-order recurrence over the <= 4
//     warps: X_w = Z_{w-1} + AW_{w-1} X_{w-1}.  Every warp publishes its end value Z in shared
//     memory, the PDE's warps meet at a named barrier, and each evaluates its own X with <= 3 DFMAs.
//     One exchange per direction, two barriers per step (the backward scan needs the corrected
//     forward values), all carry terms kept.
//   * set-up.  xDim/8 = 256 or 512 threads are needed by the cooperative set-up of Layout B
//     (setup_lu), more than the 128 that march, so it is a kernel of its own that leaves a~, g~, D, p,
//     the payoff and the chunk scalars in an HBM workspace (5*N + 3*P + 8 doubles = 165 KB per PDE at
//     4096), which the march kernel reads once.  Against 4095 time steps that round trip is 0.3 % of the
//     march's time; the workspace is sized for a chunk of the batch and reused.
// Replaces the 512-thread CTA-per-PDE kernel of Layout B for these sizes (one resident PDE per SM,
// sixteen warps in lock step at one barrier: 16 % of FP64 peak at 4096^2; this kernel: 32 %).
// The step is ~540 instructions, more than the ~512 a sub-partition's instruction cache seems to hold
// (ncu: stall_no_instruction).  A compact variant (v in TMEM, floor and e/f/Yin/Uin in shared memory,
// the chunk-pair phase as a real loop: 310 instructions) removed those stalls and lost 15-20 % to the
// extra shared-memory traffic (60 % of the smem pipe), so the unrolled form stays.
#pragma once
#include "fd1d_warp.cuh"

namespace kwfd1d {

template <int P>
struct WideSlot {
    static constexpr int N = 8 * P;
    static constexpr size_t doubles = 5 * (size_t)N + 3 * (size_t)P + 8;
    // a[N] g[N] D[N] p[N] v[N] | A[P] G[P] R0[P] | misc: [0] bmax
};

// ------------------------------------------------------------------------------------------------
// set-up: one CTA of P = N/8 threads per PDE, Layout B's code, results to the workspace
template <int P>
__global__ void __launch_bounds__(P) fd1d_wide_setup_kernel(const Fd1dBatch B, double* ws, uint32_t pde_base,
                                                             uint32_t count, int icmp)
{
    constexpr int M = 8;
    constexpr int N = 8 * P;
    extern __shared__ double smem[];
    double* xs = smem;      // [N]
    double* scr = xs + N;   // [8 * P]
    const int k = threadIdx.x;
    const int lane = k & 31;
    const int warp = k >> 5;
    const uint32_t n_pde = batch_n_pde(B);
    for (uint32_t idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const uint32_t pde = pde_base + idx;
        if (pde >= n_pde) break;  // uniform
        const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + pde) : pde;
        const PdeScalars sc = pde_scalars(load_option(B.opts + rep), B);
        double v[M], pj[M], a[M], g[M], D[M];
        setup_lu<M, P>(B, sc, icmp ? -0. : -CUDART_INF, xs, scr, v, pj, a, g, D);
        double Pp[M];
        Pp[0] = a[0];
#pragma unroll
        for (int i = 1; i < M; ++i) Pp[i] = a[i] * Pp[i - 1];
        double Q0 = g[M - 1], R0 = Pp[M - 1];
#pragma unroll
        for (int i = M - 2; i >= 0; --i) {
            Q0 = g[i] * Q0;
            R0 = fma(g[i], R0, Pp[i]);
        }
        double bmax = 0.;
#pragma unroll
        for (int i = 0; i < M; ++i) bmax = fmax(bmax, D[i] != 0. ? fabs(2. / D[i]) : 1.);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) bmax = fmax(bmax, __shfl_xor_sync(FULL, bmax, d));
        __syncthreads();  // setup_lu's scratch is free
        if (lane == 0) scr[warp] = bmax;
        __syncthreads();
        double* slot = ws + (size_t)idx * WideSlot<P>::doubles;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            slot[0 * N + k * M + i] = a[i];
            slot[1 * N + k * M + i] = g[i];
            slot[2 * N + k * M + i] = D[i];
            slot[3 * N + k * M + i] = pj[i];
            slot[4 * N + k * M + i] = v[i];
        }
        slot[5 * N + k] = Pp[M - 1];
        slot[5 * N + P + k] = Q0;
        slot[5 * N + 2 * P + k] = R0;
        if (k == 0) {
            double bm = scr[0];
            for (int w = 1; w < P / 32; ++w) bm = fmax(bm, scr[w]);
            slot[5 * N + 3 * P] = bm;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
template <int NWP>
struct WideSmem {
    static constexpr int PPC = 4 / NWP;         // PDEs per CTA
    static constexpr int N = 1024 * NWP;        // nodes per PDE tile
    // doubles: final v [PPC][N] | per-warp scan constants [4][24][32] | exchange: Zf, Zb [2 parities][4], AWf, AWb [4]
    //          | carry rows cf, cb [4 warps][4]
    static constexpr size_t bytes() { return sizeof(double) * (size_t)(PPC * N + 4 * 24 * 32 + 2 * 2 * 4 + 2 * 4 + 2 * 16 + 8); }
};

// named barrier of one PDE's warps; the id is an immediate so that only the barriers in use are reserved
template <int NTHREADS>
__device__ __forceinline__ void group_barrier(int grp)
{
    if (grp == 0)
        asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
    else
        asm volatile("bar.sync 2, %0;" ::"n"(NTHREADS) : "memory");
}

// SPLIT: as in fd1d_warp.cuh -- every chunk-pair phase a basic block of its own, a~ and g~ loaded twice.
// BS (with SPLIT): the fused FD1D-BS march -- the PDE's warps march the chain as given, price it into B.prices, re-create the
// payoff and march the European copy (no floor, no compare) with a~, g~, D still in tensor memory into B.prices_eu; a chain given
// as European is marched once and priced into both arrays (reference src/Pricer/kwFd1d_BlackScholes.cpp:15-43).
template <int NWP, int MINB, bool ICMP, bool SPLIT = false, bool BS = false>
__global__ void __launch_bounds__(128, MINB) fd1d_wide_kernel(const Fd1dBatch B, const double* __restrict__ ws,
                                                               uint32_t pde_base, uint32_t count)
{
    static_assert(NWP == 2 || NWP == 4, "two or four warps per PDE");
    static_assert(!BS || SPLIT, "the fused FD1D-BS march exists in the rotated split form only");
    constexpr int NCH = 4;
    constexpr int NODES = 32;                 // per lane
    constexpr int PPC = 4 / NWP;
    constexpr int N = 1024 * NWP;
    constexpr int P = N / 8;                  // chunks per PDE
    using Slot = WideSlot<P>;

    extern __shared__ double smem[];
    double* vfin_all = smem;                       // [PPC][N]
    double* wconst = vfin_all + PPC * N;           // [4][24][32]
    double* zf = wconst + 4 * 24 * 32;             // [2][4]
    double* zb = zf + 8;                           // [2][4]
    double* awf = zb + 8;                          // [4]
    double* awb = awf + 4;                         // [4]
    double* cfr = awb + 4;                         // [4 warps][4]: X_w  = sum_w' cf[w][w'] Zf[w']
    double* cbr = cfr + 16;                        // [4 warps][4]: Xb_w = sum_w' cb[w][w'] Zb[w']

    const int k = threadIdx.x;
    const int lane = k & 31;
    const int warp = k >> 5;
    const int grp = warp / NWP;      // which PDE of the CTA
    const int wq = warp % NWP;       // warp inside the PDE
    const int w0 = grp * NWP;        // first warp of the PDE
    const int xDim = B.xDim;
    const int nsteps = B.tDim - 1;

    __shared__ uint32_t s_taddr;
    if (warp == 0) tmem::alloc<256>(smem_addr(&s_taddr));
    tmem::fence_before();
    __syncthreads();
    tmem::fence_after();
    const uint32_t tbase = s_taddr + ((uint32_t)(warp & 3) << 21);
    const uint32_t tbase2 = SPLIT ? tbase + B.opq_zero : tbase;
    constexpr uint32_t T_A = 0, T_G = 64, T_D = 128, T_P = 192;

    const uint32_t n_pde = batch_n_pde(B);
    const uint32_t n_iter = (count + PPC - 1) / PPC;
    for (uint32_t it = blockIdx.x; it < n_iter; it += gridDim.x) {
        const uint32_t idx = it * PPC + grp;
        const uint32_t my_pde = pde_base + idx;
        const bool have = idx < count && my_pde < n_pde;  // uniform over the PDE's warps
        double* vfin = vfin_all + grp * N;
        if (have) {
            const double* slot = ws + (size_t)idx * Slot::doubles;
            double vr[NODES];
            double Ac[NCH], Gc[NCH];
            double* wc = wconst + warp * (24 * 32) + lane;
            // ---------------- this lane's four chunks out of the workspace -----------------------
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int ch = (wq * 32 + lane) * NCH + c;
                double t8[8];
#pragma unroll
                for (int arr = 0; arr < 4; ++arr) {
                    const double2* src = reinterpret_cast<const double2*>(slot + (size_t)arr * N + ch * 8);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double2 d = __ldg(src + i);
                        t8[2 * i] = d.x;
                        t8[2 * i + 1] = d.y;
                    }
                    tmem::st8(tbase + 64 * arr + 16 * c, t8);
                }
                const double2* srcv = reinterpret_cast<const double2*>(slot + (size_t)4 * N + ch * 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double2 d = __ldg(srcv + i);
                    vr[8 * c + 2 * i] = d.x;
                    vr[8 * c + 2 * i + 1] = d.y;
                }
                Ac[c] = __ldg(slot + 5 * N + ch);
                Gc[c] = __ldg(slot + 5 * N + P + ch);
                wc[(0 + c) * 32] = Ac[c];
                wc[(4 + c) * 32] = Gc[c];
                wc[(8 + c) * 32] = __ldg(slot + 5 * N + 2 * P + ch);
            }
            tmem::wait_st();
            const double bm = __ldg(slot + 5 * N + 3 * P);
            // ---------------- lane aggregates, in-warp scan multipliers, boundary responses ------
            double AL = Ac[0], GL = Gc[0];
#pragma unroll
            for (int c = 1; c < NCH; ++c) {
                AL *= Ac[c];
                GL *= Gc[c];
            }
            double AfL[5], GbL[5], PWfex, PWbex;
            {
                double A = AL;
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int s = 1 << d;
                    const double o = __shfl_up_sync(FULL, A, s);
                    AfL[d] = lane >= s ? A : 0.;
                    if (lane >= s) A *= o;
                }
                // A = product of the lane multipliers warp start .. this lane
                const double ex = __shfl_up_sync(FULL, A, 1);
                PWfex = lane ? ex : 1.;
                if (lane == 31) awf[warp] = A;
                double G = GL;
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int s = 1 << d;
                    const double o = __shfl_down_sync(FULL, G, s);
                    GbL[d] = lane < 32 - s ? G : 0.;
                    if (lane < 32 - s) G *= o;
                }
                const double exb = __shfl_down_sync(FULL, G, 1);
                PWbex = lane < 31 ? exb : 1.;
                if (lane == 0) awb[warp] = G;
            }
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                wc[(12 + d) * 32] = AfL[d];
                wc[(17 + d) * 32] = GbL[d];
            }
            wc[22 * 32] = PWfex;
            wc[23 * 32] = PWbex;
            // ---------------- scan levels that carry anything (in-warp; cross-warp terms are all kept) ----
            int levels;
            {
                const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + my_pde) : my_pde;
                const PdeScalars sc = pde_scalars(load_option(B.opts + rep), B);
                const double tol = 0x1p-56 / (bm * (double)B.tDim);
                const int l_abs = wq * 32 + lane;
                const double x_here = fmax(0., x_node(sc, B.density, min(l_abs * NODES, xDim - 1)));
                int lv = 0;
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int src = min((l_abs + (1 << d)) * NODES + NODES - 1, xDim - 1);
                    const double growth = sc.put ? 1. : exp(fmax(0., x_node(sc, B.density, src)) - x_here);
                    const bool bad = !(fabs(AfL[d]) <= tol) || !(fabs(GbL[d]) * growth <= tol);
                    if (__any_sync(FULL, bad)) lv = d + 1;
                }
                if (B.max_mode <= 1) lv = 5;
                levels = lv < 1 ? 1 : lv;
                // the PDE's warps must run the same number of levels: they share the barriers, not the
                // code path, but a common value keeps the histogram per PDE
            }
            group_barrier<32 * NWP>(grp);  // awf / awb visible to the PDE's warps
            if (lane == 0) {
                // X_w = Z_{w-1} + AW_{w-1} X_{w-1} unrolled into one row of weights per warp (0 outside the PDE
                // or on the wrong side), so that the march evaluates its carry as a branch-free dot product
                double c = 1.;
                for (int w = 3; w >= 0; --w) {
                    double val = 0.;
                    if (w >= w0 && w < warp) {
                        val = c;
                        c *= awf[w];
                    }
                    cfr[warp * 4 + w] = val;
                }
                c = 1.;
                for (int w = 0; w < 4; ++w) {
                    double val = 0.;
                    if (w > warp && w < w0 + NWP) {
                        val = c;
                        c *= awb[w];
                    }
                    cbr[warp * 4 + w] = val;
                }
            }
            __syncwarp();
            const uint32_t a_wc = smem_addr(wc);
            auto K = [&](int idx2) { return lds_f64(a_wc + idx2 * 256); };
            const uint32_t a_zf = smem_addr(zf), a_zb = smem_addr(zb);
            const uint32_t a_cf = smem_addr(cfr + warp * 4), a_cb = smem_addr(cbr + warp * 4);
            const uint32_t a_myzf = a_zf + warp * 8, a_myzb = a_zb + warp * 8;
            const bool last_lane = lane == 31, first_lane = lane == 0;

            auto march = [&](auto lev_c) {
                constexpr int LEV = decltype(lev_c)::value;
                double e[NCH], f[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    double a8[8], g8[8];
                    tmem::ld8(tbase + T_A + 16 * c, a8);
                    tmem::ld8(tbase + T_G + 16 * c, g8);
                    tmem::wait_ld_dep(a8);
                    tmem::wait_ld_dep(g8);
                    double y[8];
                    y[0] = vr[8 * c];
#pragma unroll
                    for (int i = 1; i < 8; ++i) y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                    e[c] = y[7];
                    double u = y[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) u = fma(g8[i], u, y[i]);
                    f[c] = u;
                }
                for (int step = 0; step < nsteps; ++step) {
                    const uint32_t par = (uint32_t)(step & 1) * 32;  // bytes: 4 doubles per parity
                    // ---- forward: lane aggregate, in-warp scan, cross-warp carry, chunk-entry values
                    double S = e[0];
#pragma unroll
                    for (int c = 1; c < NCH; ++c) S = fma(K(c), S, e[c]);
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const double o = __shfl_up_sync(FULL, S, 1 << d);
                        S = fma(K(12 + d), o, S);
                    }
                    if (last_lane) sts_f64(a_myzf + par, S);
                    double Sm1 = __shfl_up_sync(FULL, S, 1);
                    if (first_lane) Sm1 = 0.;
                    group_barrier<32 * NWP>(grp);
                    double X = 0.;
#pragma unroll
                    for (int w = 0; w < NWP; ++w)  // the PDE's own slots only: another PDE's value times 0 may be NaN
                        X = fma(lds_f64(a_cf + (w0 + w) * 8), lds_f64(a_zf + par + (w0 + w) * 8), X);
                    double Yin[NCH];
                    Yin[0] = fma(K(22), X, Sm1);
#pragma unroll
                    for (int c = 1; c < NCH; ++c) Yin[c] = fma(K(c - 1), Yin[c - 1], e[c - 1]);
                    // ---- backward
#pragma unroll
                    for (int c = 0; c < NCH; ++c) f[c] = fma(K(8 + c), Yin[c], f[c]);
                    double T = f[NCH - 1];
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) T = fma(K(4 + c), T, f[c]);
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const double o = __shfl_down_sync(FULL, T, 1 << d);
                        T = fma(K(17 + d), o, T);
                    }
                    if (first_lane) sts_f64(a_myzb + par, T);
                    double Tp1 = __shfl_down_sync(FULL, T, 1);
                    if (last_lane) Tp1 = 0.;
                    group_barrier<32 * NWP>(grp);
                    double Xb = 0.;
#pragma unroll
                    for (int w = 0; w < NWP; ++w)
                        Xb = fma(lds_f64(a_cb + (w0 + w) * 8), lds_f64(a_zb + par + (w0 + w) * 8), Xb);
                    double Uin[NCH];
                    Uin[NCH - 1] = fma(K(23), Xb, Tp1);
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) Uin[c] = fma(K(4 + c + 1), Uin[c + 1], f[c + 1]);
                    // ---- chunk pairs: true sweeps, projection, next step's local sweeps
#pragma unroll
                    for (int h = 0; h < NCH; h += 2) {
                        if (SPLIT && !(step < B.opq_lim[h >> 1])) continue;  // never taken: basic-block boundary
                        const int cA = h, cB = h + 1;
                        double aA[8], aB[8], gA[8], gB[8], dA[8], dB[8], pA[8], pB[8];
                        tmem::ld8(tbase + T_A + 16 * cA, aA);
                        tmem::ld8(tbase + T_A + 16 * cB, aB);
                        tmem::wait_ld_dep(aA);
                        tmem::wait_ld_dep(aB);
                        tmem::ld8(tbase + T_G + 16 * cA, gA);
                        tmem::ld8(tbase + T_G + 16 * cB, gB);
                        tmem::ld8(tbase + T_D + 16 * cA, dA);
                        tmem::ld8(tbase + T_D + 16 * cB, dB);
                        tmem::ld8(tbase + T_P + 16 * cA, pA);
                        tmem::ld8(tbase + T_P + 16 * cB, pB);
                        double yA[8], yB[8];
                        yA[0] = fma(aA[0], Yin[cA], vr[8 * cA]);
                        yB[0] = fma(aB[0], Yin[cB], vr[8 * cB]);
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                            yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                        }
                        tmem::wait_ld_dep(gA);
                        tmem::wait_ld_dep(gB);
                        tmem::wait_ld_dep(dA);
                        tmem::wait_ld_dep(dB);
                        tmem::wait_ld_dep(pA);
                        tmem::wait_ld_dep(pB);
                        double uA = Uin[cA], uB = Uin[cB];
#pragma unroll
                        for (int i = 7; i >= 0; --i) {
                            uA = fma(gA[i], uA, yA[i]);
                            uB = fma(gB[i], uB, yB[i]);
                            const double rA = fma(dA[i], uA, -vr[8 * cA + i]);
                            const double rB = fma(dB[i], uB, -vr[8 * cB + i]);
                            vr[8 * cA + i] = ICMP ? max_like_icmp(rA, pA[i]) : max_like_std(rA, pA[i]);
                            vr[8 * cB + i] = ICMP ? max_like_icmp(rB, pB[i]) : max_like_std(rB, pB[i]);
                        }
                        if constexpr (SPLIT) {
                            tmem::ld8(tbase2 + T_A + 16 * cA, aA);
                            tmem::ld8(tbase2 + T_A + 16 * cB, aB);
                            tmem::ld8(tbase2 + T_G + 16 * cA, gA);
                            tmem::ld8(tbase2 + T_G + 16 * cB, gB);
                            tmem::wait_ld_dep(aA, aB);
                            tmem::wait_ld_dep(gA, gB);
                        }
                        yA[0] = vr[8 * cA];
                        yB[0] = vr[8 * cB];
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                            yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                        }
                        e[cA] = yA[7];
                        e[cB] = yB[7];
                        uA = yA[7];
                        uB = yB[7];
#pragma unroll
                        for (int i = 6; i >= 0; --i) {
                            uA = fma(gA[i], uA, yA[i]);
                            uB = fma(gB[i], uB, yB[i]);
                        }
                        f[cA] = uA;
                        f[cB] = uB;
                    }
                }
            };
            // SPLIT: the rotated form of fd1d_iw.cuh -- scan constants in registers, every chunk-pair phase a basic
            // block, a~ / g~ loaded twice, and the NEXT step's scans (with their two named barriers) issued behind the
            // last pair's local sweeps with pair 0's forward sweeps behind them, so that shuffle, shared-memory and
            // barrier latency hides behind sweeps of the same basic block.
            auto march_rot = [&](auto lev_c, auto euro_c) {
                constexpr int LEV = decltype(lev_c)::value;
                constexpr bool EURO = decltype(euro_c)::value;  // European copy: no floor, no compare
                double kA[NCH], kG[NCH], kR[NCH], kAf[LEV], kGb[LEV];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    kA[c] = K(c);
                    kG[c] = K(4 + c);
                    kR[c] = K(8 + c);
                }
#pragma unroll
                for (int d = 0; d < LEV; ++d) {
                    kAf[d] = K(12 + d);
                    kGb[d] = K(17 + d);
                }
                const double kPf = K(22), kPb = K(23);
                double e[NCH], f[NCH], Yin[NCH], Uin[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    double a8[8], g8[8];
                    tmem::ld8(tbase + T_A + 16 * c, a8);
                    tmem::ld8(tbase + T_G + 16 * c, g8);
                    tmem::wait_ld_dep(a8, g8);
                    double y[8];
                    y[0] = vr[8 * c];
#pragma unroll
                    for (int i = 1; i < 8; ++i) y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                    e[c] = y[7];
                    double u = y[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) u = fma(g8[i], u, y[i]);
                    f[c] = u;
                }
                uint32_t par = 0;  // bytes: 4 doubles per parity, flipped by every scan
                auto scan_fwd = [&]() {
                    double S = e[0];
#pragma unroll
                    for (int c = 1; c < NCH; ++c) S = fma(kA[c], S, e[c]);
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const double o = __shfl_up_sync(FULL, S, 1 << d);
                        S = fma(kAf[d], o, S);
                    }
                    if (last_lane) sts_f64(a_myzf + par, S);
                    double Sm1 = __shfl_up_sync(FULL, S, 1);
                    if (first_lane) Sm1 = 0.;
                    group_barrier<32 * NWP>(grp);
                    double X = 0.;
#pragma unroll
                    for (int w = 0; w < NWP; ++w)  // the PDE's own slots only: another PDE's value times 0 may be NaN
                        X = fma(lds_f64(a_cf + (w0 + w) * 8), lds_f64(a_zf + par + (w0 + w) * 8), X);
                    Yin[0] = fma(kPf, X, Sm1);
#pragma unroll
                    for (int c = 1; c < NCH; ++c) Yin[c] = fma(kA[c - 1], Yin[c - 1], e[c - 1]);
                };
                auto scan_bwd = [&]() {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) f[c] = fma(kR[c], Yin[c], f[c]);
                    double T = f[NCH - 1];
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) T = fma(kG[c], T, f[c]);
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const double o = __shfl_down_sync(FULL, T, 1 << d);
                        T = fma(kGb[d], o, T);
                    }
                    if (first_lane) sts_f64(a_myzb + par, T);
                    double Tp1 = __shfl_down_sync(FULL, T, 1);
                    if (last_lane) Tp1 = 0.;
                    group_barrier<32 * NWP>(grp);
                    double Xb = 0.;
#pragma unroll
                    for (int w = 0; w < NWP; ++w)
                        Xb = fma(lds_f64(a_cb + (w0 + w) * 8), lds_f64(a_zb + par + (w0 + w) * 8), Xb);
                    Uin[NCH - 1] = fma(kPb, Xb, Tp1);
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) Uin[c] = fma(kG[c + 1], Uin[c + 1], f[c + 1]);
                    par ^= 32u;
                };
                auto fwd_load = [&](int cA, double (&aA)[8], double (&aB)[8]) {
                    tmem::ld8(tbase + T_A + 16 * cA, aA);
                    tmem::ld8(tbase + T_A + 16 * (cA + 1), aB);
                    tmem::hot_wait(aA, aB);
                };
                auto fwd_sweep = [&](int cA, const double (&aA)[8], const double (&aB)[8], double (&yA)[8], double (&yB)[8]) {
                    const int cB = cA + 1;
                    yA[0] = fma(aA[0], Yin[cA], vr[8 * cA]);
                    yB[0] = fma(aB[0], Yin[cB], vr[8 * cB]);
#pragma unroll
                    for (int i = 1; i < 8; ++i) {
                        yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                        yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                    }
                };
                auto back_pair = [&](int cA, double (&yA)[8], double (&yB)[8]) {
                    const int cB = cA + 1;
                    double gA[8], gB[8], dA[8], dB[8], pA[8], pB[8];
                    tmem::ld8(tbase + T_G + 16 * cA, gA);
                    tmem::ld8(tbase + T_G + 16 * cB, gB);
                    tmem::ld8(tbase + T_D + 16 * cA, dA);
                    tmem::ld8(tbase + T_D + 16 * cB, dB);
                    if constexpr (!EURO) {
                        tmem::ld8(tbase + T_P + 16 * cA, pA);
                        tmem::ld8(tbase + T_P + 16 * cB, pB);
                        tmem::hot_wait(gA, gB, dA);
                        tmem::hot_wait(dB, pA, pB);
                    } else {
                        tmem::hot_wait(gA, gB);
                        tmem::hot_wait(dA, dB);
                    }
                    double uA = Uin[cA], uB = Uin[cB];
#pragma unroll
                    for (int i = 7; i >= 0; --i) {
                        uA = fma(gA[i], uA, yA[i]);
                        uB = fma(gB[i], uB, yB[i]);
                        const double rA = fma(dA[i], uA, -vr[8 * cA + i]);
                        const double rB = fma(dB[i], uB, -vr[8 * cB + i]);
                        if constexpr (EURO) {
                            vr[8 * cA + i] = rA;
                            vr[8 * cB + i] = rB;
                        } else {
                            vr[8 * cA + i] = ICMP ? max_like_icmp(rA, pA[i]) : max_like_std(rA, pA[i]);
                            vr[8 * cB + i] = ICMP ? max_like_icmp(rB, pB[i]) : max_like_std(rB, pB[i]);
                        }
                    }
                    double aA[8], aB[8];
                    tmem::ld8(tbase2 + T_A + 16 * cA, aA);
                    tmem::ld8(tbase2 + T_A + 16 * cB, aB);
                    tmem::ld8(tbase2 + T_G + 16 * cA, gA);
                    tmem::ld8(tbase2 + T_G + 16 * cB, gB);
                    tmem::hot_wait(aA, aB);
                    tmem::hot_wait(gA, gB);
                    yA[0] = vr[8 * cA];
                    yB[0] = vr[8 * cB];
#pragma unroll
                    for (int i = 1; i < 8; ++i) {
                        yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                        yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                    }
                    e[cA] = yA[7];
                    e[cB] = yB[7];
                    uA = yA[7];
                    uB = yB[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) {
                        uA = fma(gA[i], uA, yA[i]);
                        uB = fma(gB[i], uB, yB[i]);
                    }
                    f[cA] = uA;
                    f[cB] = uB;
                };
                // pair 0's a~ is loaded between the two halves of the scan: its forward sweeps need only Yin, so they
                // fill the wait at the second named barrier
                double y0A[8], y0B[8];
                {
                    double aA[8], aB[8];
                    scan_fwd();
                    fwd_load(0, aA, aB);
                    scan_bwd();
                    fwd_sweep(0, aA, aB, y0A, y0B);
                }
                for (int step = 0; step < nsteps; ++step) {
                    if (step < B.opq_lim[0]) back_pair(0, y0A, y0B);  // always true: basic-block boundary
                    if (step < B.opq_lim[1]) {
                        double yA[8], yB[8], aA[8], aB[8];
                        fwd_load(2, aA, aB);
                        fwd_sweep(2, aA, aB, yA, yB);
                        back_pair(2, yA, yB);
                        scan_fwd();
                        fwd_load(0, aA, aB);
                        scan_bwd();
                        fwd_sweep(0, aA, aB, y0A, y0B);  // after the last step: computed and dropped
                    }
                }
            };
            // the PDE's warps share barriers, so they must agree on nothing but the step count; each picks
            // its own number of in-warp levels
            auto march_levels = [&](auto euro_c) {
                if constexpr (SPLIT) {
                    switch (levels) {
                        case 1: march_rot(std::integral_constant<int, 1>{}, euro_c); break;
                        case 2: march_rot(std::integral_constant<int, 2>{}, euro_c); break;
                        case 3: march_rot(std::integral_constant<int, 3>{}, euro_c); break;
                        case 4: march_rot(std::integral_constant<int, 4>{}, euro_c); break;
                        default: march_rot(std::integral_constant<int, 5>{}, euro_c); break;
                    }
                } else {
                    switch (levels) {
                        case 1: march(std::integral_constant<int, 1>{}); break;
                        case 2: march(std::integral_constant<int, 2>{}); break;
                        case 3: march(std::integral_constant<int, 3>{}); break;
                        case 4: march(std::integral_constant<int, 4>{}); break;
                        default: march(std::integral_constant<int, 5>{}); break;
                    }
                }
            };
            const uint32_t rep_e = B.pde_rep ? __ldg(B.pde_rep + my_pde) : my_pde;
            const PdeScalars sc_e = pde_scalars(load_option(B.opts + rep_e), B);
            // ---------------- epilogue: every option of the chain, interpolated by the PDE's warps ------------------
            auto emit = [&](double* out) {
#pragma unroll
                for (int i = 0; i < NODES; ++i) vfin[(wq * 32 + lane) * NODES + i] = vr[i];
                group_barrier<32 * NWP>(grp);
                Fd1dBatch Bo = B;
                Bo.prices = out;
                uint32_t q0, q1;
                chain_range(B, my_pde, q0, q1);
                for (uint32_t q = q0 + wq * 32 + lane; q < q1; q += 32 * NWP) {
                    const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                    price_option(Bo, oi, [&](int j) { return x_node(sc_e, B.density, j); }, [&](int j) { return vfin[j]; });
                }
                group_barrier<32 * NWP>(grp);  // vfin is rewritten by the next emit
            };
            march_levels(std::false_type{});
            if (lane == 0 && wq == 0) {
                const int bucket = B.max_mode == 0 ? 0 : (levels == 5 ? 1 : 6 - levels);
                atomicAdd(&B.status[2 + bucket], 1u);
            }
            emit(B.prices);
            if constexpr (BS) {
                if (sc_e.american) {  // uniform over the PDE's warps
                    // the European copy: payoff again (src/Pricer/kwFd1d.cpp:127-139, as the set-up kernel computed it)
#pragma unroll 1
                    for (int c = 0; c < NCH; ++c) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int j = (wq * 32 + lane) * NODES + 8 * c + i;
                            vfin[j] = j < xDim ? payoff_node(sc_e.put, x_node(sc_e, B.density, j)) : 0.;
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < NODES; ++i) vr[i] = vfin[(wq * 32 + lane) * NODES + i];
                    __syncwarp();
                    march_levels(std::true_type{});
                }
                emit(B.prices_eu);  // a chain given as European: one march, both arrays
            }
        }
        __syncthreads();  // vfin, the exchange slots and the TMEM arrays are rewritten by the next PDE
    }
    tmem::fence_before();
    __syncthreads();
    if (warp == 0) tmem::dealloc<256>(s_taddr);
}

}  // namespace kwfd1d
