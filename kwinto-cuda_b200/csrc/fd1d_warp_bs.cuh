// fd1d_warp_bs.cuh -- Layout W, fused march for the control-variate pricer "FD1D-BS".
//
// Fd1d_BlackScholes_Pricer::price (reference src/Pricer/kwFd1d_BlackScholes.cpp:15-43) solves every
// chain twice: as given (e = 1: with the early-exercise projection) and as a European copy (e = 0),
// and returns FD_given + (BS_european - FD_european).  Both solves share the grid, the matrix
// B = 1 - dt/2 A and therefore every coefficient of the hoisted LU (a~, g~, D); only the projection
// differs.  This kernel marches the two solution vectors of a chain side by side in one warp:
//   * the "pair" of interleaved dependent chains of fd1d_warp.cuh is (given, European) on the SAME
//     chunk instead of two chunks of one solution: one tcgen05.ld of a~, g~, D feeds both, and every
//     coefficient register is the shared source operand of two consecutive DFMAs -- the one place in
//     this scheme where the operand-fetch bound of DESIGN.md section 5 (three distinct register
//     sources = 3 issue slots per DFMA) relaxes to 2.5 slots;
//   * v_given stays in registers, v_european lives in tensor memory next to a~, g~, D (read for the
//     sweeps, written back after every step), the projection floor comes from shared memory two
//     nodes at a time (the arrangement of fd1d_warp2_kernel);
//   * the scan constants are read once per step for both solutions.
// The epilogue interpolates both solutions for every option of the chain: prices[] (as given) and
// prices_eu[] (European); the caller adds the closed form (bs_combine_kernel, capi.cu).
// Set-up, truncation proof, launch geometry: fd1d_warp2_kernel's.  512 < xDim <= 1024.
//
// MEASURED (profiles/r1_as_*, r1_at_*): prices bit-identical to the two-solve path, but 82.6 ms against
// 2 x 29.2 ms for 32768 chains at 1024^2.  The fully unrolled step is ~800 instructions (12.7 KB); ncu
// reports stall_no_instruction = 3.8 cycles per issued instruction (0.06 for the 444-instruction step of
// fd1d_warp_kernel): the loop no longer fits the instruction cache, and ptxas schedules only 30 of the 340
// DFMAs next to their partner (.reuse).  Hence opt-in (FD1D.GPU.BS_FUSED = 2), not the default; a version
// that fits needs a rolled chunk loop with both solutions in tensor memory (DESIGN.md "Next").
#pragma once
#include "fd1d_warp.cuh"

namespace kwfd1d {

// Fd1d::value for both solutions of one option (src/Math/kwFd1d.cpp:139-158; same search, same formula)
template <class XS, class VA, class VE>
__device__ __forceinline__ void price_option2(const Fd1dBatch& B, uint32_t oi, XS xs, VA va, VE ve)
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
        B.prices_eu[oi] = CUDART_NAN;
        atomicAdd(&B.status[0], 1u);
        atomicMin(&B.status[1], oi);
        return;
    }
    const double x1 = xs(lo), x0 = xs(lo - 1);
    const double wl = x1 - xq, wr = xq - x0, dx = x1 - x0;
    const double na = __dadd_rn(__dmul_rn(wl, va(lo - 1)), __dmul_rn(wr, va(lo)));
    const double ne = __dadd_rn(__dmul_rn(wl, ve(lo - 1)), __dmul_rn(wr, ve(lo)));
    B.prices[oi] = __dmul_rn(o.k, na / dx);
    B.prices_eu[oi] = __dmul_rn(o.k, ne / dx);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) fd1d_warp_bs_kernel(const Fd1dBatch B)
{
    constexpr int NCH = 4;
    using L = Warp2Smem<NCH>;
    constexpr int N = L::N;
    constexpr int P = L::P;
    constexpr int M = 8;
    constexpr int NODES = 8 * NCH;

    extern __shared__ double smem[];
    double* xs = smem;            // [N]
    double* st = xs + N;          // [5][N] set-up stage; after the set-up: the final European solutions [4][N]
    double* scr = st + 4 * N;     // setup_lu scratch = the v slot of the stage (free until staging)
    double* ps = st + 5 * N;      // [4][N] floors, [chunk][node pair][lane] double2; after the march: final v as given
    double* st_A = ps + 4 * N;    // [P]
    double* st_G = st_A + P;
    double* st_R0 = st_G + P;
    double* misc = st_R0 + P;     // [0..3] per-warp bmax of the PDE being set up
    double* wconst = misc + 16;   // [4][22][32]

    const int k = threadIdx.x;
    const int lane = k & 31;
    const int warp = k >> 5;
    const int xDim = B.xDim;
    const int nsteps = B.tDim - 1;

    __shared__ uint32_t s_taddr;
    if (warp == 0) tmem::alloc<64 * NCH>(smem_addr(&s_taddr));
    tmem::fence_before();
    __syncthreads();
    tmem::fence_after();
    const uint32_t tbase = s_taddr + ((uint32_t)(warp & 3) << 21);
    constexpr uint32_t T_A = 0, T_G = 16 * NCH, T_D = 32 * NCH, T_VE = 48 * NCH;

    const uint32_t n_pde = batch_n_pde(B);
    const uint32_t n_grp = (n_pde + 3) / 4;
    for (uint32_t grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        int levels = 5;
        double vr[NODES];  // the solution as given (projected when e = 1), this lane's 32 nodes
        double* wc = wconst + warp * (22 * 32) + lane;
        double* myp = ps + warp * N;

        // ---------------- set-up, one PDE at a time, all 128 threads ---------------------------
        for (int q = 0; q < 4; ++q) {
            const uint32_t pde = 4 * grp + q;
            if (pde >= n_pde) break;  // uniform across the CTA
            const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + pde) : pde;
            const kw_option opt = load_option(B.opts + rep);
            const PdeScalars sc = pde_scalars(opt, B);
            {
                double v[M], pj[M], a[M], g[M], D[M];
                setup_lu<M, P>(B, sc, -CUDART_INF, xs, scr, v, pj, a, g, D);
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
                __syncthreads();  // setup_lu's scratch (= the v slot) is free
                if (lane == 0) misc[warp] = bmax;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    st[0 * N + k * M + i] = a[i];
                    st[1 * N + k * M + i] = g[i];
                    st[2 * N + k * M + i] = D[i];
                    st[3 * N + k * M + i] = pj[i];
                    st[4 * N + k * M + i] = v[i];
                }
                st_A[k] = Pp[M - 1];
                st_G[k] = Q0;
                st_R0[k] = R0;
            }
            __syncthreads();
            if (warp == q) {
                // the owner pulls its lane's chunks: a~, g~, D and the payoff (European start) into TMEM,
                // the payoff into registers as well, the floor into its shared array
                double Ac[NCH], Gc[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int ch = lane * NCH + c;
                    double t8[8];
#pragma unroll
                    for (int arr = 0; arr < 4; ++arr) {
                        const int from = arr == 3 ? 4 : arr;  // TMEM slot 3 holds the European solution
#pragma unroll
                        for (int i = 0; i < 8; ++i) t8[i] = st[from * N + ch * 8 + i];
                        tmem::st8(tbase + 16 * NCH * arr + 16 * c, t8);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) vr[8 * c + i] = t8[i];
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2)
                        reinterpret_cast<double2*>(myp)[(c * 4 + i2) * 32 + lane] =
                            make_double2(st[3 * N + ch * 8 + 2 * i2], st[3 * N + ch * 8 + 2 * i2 + 1]);
                    Ac[c] = st_A[ch];
                    Gc[c] = st_G[ch];
                    wc[(0 + c) * 32] = Ac[c];
                    wc[(4 + c) * 32] = Gc[c];
                    wc[(8 + c) * 32] = st_R0[ch];
                }
                tmem::wait_st();
                const double bm = fmax(fmax(misc[0], misc[1]), fmax(misc[2], misc[3]));
                double AL = Ac[0], GL = Gc[0];
#pragma unroll
                for (int c = 1; c < NCH; ++c) {
                    AL *= Ac[c];
                    GL *= Gc[c];
                }
                double AfL[5], GbL[5];
                {
                    double A = AL;
#pragma unroll
                    for (int d = 0; d < 5; ++d) {
                        const int s = 1 << d;
                        const double o = __shfl_up_sync(FULL, A, s);
                        AfL[d] = lane >= s ? A : 0.;
                        if (lane >= s) A *= o;
                    }
                    double G = GL;
#pragma unroll
                    for (int d = 0; d < 5; ++d) {
                        const int s = 1 << d;
                        const double o = __shfl_down_sync(FULL, G, s);
                        GbL[d] = lane < 32 - s ? G : 0.;
                        if (lane < 32 - s) G *= o;
                    }
                }
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    wc[(12 + d) * 32] = AfL[d];
                    wc[(17 + d) * 32] = GbL[d];
                }
                // how many levels carry anything (DESIGN.md "Truncation"); the bound holds for both
                // solutions (the European one is dominated by the projected one)
                const double tol = 0x1p-56 / (bm * (double)B.tDim);
                const double x_here = fmax(0., xs[min(lane * NODES, xDim - 1)]);
                int lv = 0;
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int src = min((lane + (1 << d)) * NODES + NODES - 1, xDim - 1);
                    const double growth = sc.put ? 1. : exp(fmax(0., xs[src]) - x_here);
                    const bool bad = !(fabs(AfL[d]) <= tol) || !(fabs(GbL[d]) * growth <= tol);
                    if (__any_sync(FULL, bad)) lv = d + 1;
                }
                if (B.max_mode <= 1) lv = 5;
                levels = lv < 1 ? 1 : lv;
            }
            __syncthreads();
        }

        const uint32_t my_pde = 4 * grp + warp;
        const bool have = my_pde < n_pde;  // warp-uniform
        if (have) {
            const uint32_t a_wc = smem_addr(wc);
            const uint32_t a_p = smem_addr(myp) + lane * 16;
            auto K = [&](int idx) { return lds_f64(a_wc + idx * 256); };

            auto march = [&](auto lev_c) {
                constexpr int LEV = decltype(lev_c)::value;
                double eA[NCH], fA[NCH], eE[NCH], fE[NCH];
                // first local sweeps: both solutions start from the payoff
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
                    eA[c] = eE[c] = y[7];
                    double u = y[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) u = fma(g8[i], u, y[i]);
                    fA[c] = fE[c] = u;
                }
                for (int step = 0; step < nsteps; ++step) {
                    // ---- forward: lane aggregates, scans over lanes, chunk-entry values (both solutions)
                    double SA = eA[0], SE = eE[0];
#pragma unroll
                    for (int c = 1; c < NCH; ++c) {
                        const double kc = K(c);
                        SA = fma(kc, SA, eA[c]);
                        SE = fma(kc, SE, eE[c]);
                    }
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const double kd = K(12 + d);
                        const double oA = __shfl_up_sync(FULL, SA, 1 << d);
                        const double oE = __shfl_up_sync(FULL, SE, 1 << d);
                        SA = fma(kd, oA, SA);
                        SE = fma(kd, oE, SE);
                    }
                    double YA[NCH], YE[NCH];
                    {
                        const double oA = __shfl_up_sync(FULL, SA, 1);
                        const double oE = __shfl_up_sync(FULL, SE, 1);
                        YA[0] = lane ? oA : 0.;
                        YE[0] = lane ? oE : 0.;
                    }
#pragma unroll
                    for (int c = 1; c < NCH; ++c) {
                        const double kc = K(c - 1);
                        YA[c] = fma(kc, YA[c - 1], eA[c - 1]);
                        YE[c] = fma(kc, YE[c - 1], eE[c - 1]);
                    }
                    // ---- backward
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const double kc = K(8 + c);
                        fA[c] = fma(kc, YA[c], fA[c]);
                        fE[c] = fma(kc, YE[c], fE[c]);
                    }
                    double TA = fA[NCH - 1], TE = fE[NCH - 1];
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) {
                        const double kc = K(4 + c);
                        TA = fma(kc, TA, fA[c]);
                        TE = fma(kc, TE, fE[c]);
                    }
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const double kd = K(17 + d);
                        const double oA = __shfl_down_sync(FULL, TA, 1 << d);
                        const double oE = __shfl_down_sync(FULL, TE, 1 << d);
                        TA = fma(kd, oA, TA);
                        TE = fma(kd, oE, TE);
                    }
                    double UA[NCH], UE[NCH];
                    {
                        const double oA = __shfl_down_sync(FULL, TA, 1);
                        const double oE = __shfl_down_sync(FULL, TE, 1);
                        UA[NCH - 1] = lane < 31 ? oA : 0.;
                        UE[NCH - 1] = lane < 31 ? oE : 0.;
                    }
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) {
                        const double kc = K(4 + c + 1);
                        UA[c] = fma(kc, UA[c + 1], fA[c + 1]);
                        UE[c] = fma(kc, UE[c + 1], fE[c + 1]);
                    }
                    tmem::wait_st();  // last step's European values are in place
                    // ---- per chunk: true sweeps of both solutions from (Yin, Uin), projection of the
                    //      one as given, European values back to TMEM, next step's local sweeps
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        double a8[8], w8[8], g8[8], d8[8];
                        tmem::ld8(tbase + T_A + 16 * c, a8);
                        tmem::ld8(tbase + T_VE + 16 * c, w8);
                        tmem::wait_ld_dep(a8);
                        tmem::wait_ld_dep(w8);
                        tmem::ld8(tbase + T_G + 16 * c, g8);
                        tmem::ld8(tbase + T_D + 16 * c, d8);
                        double yA[8], yE[8];
                        yA[0] = fma(a8[0], YA[c], vr[8 * c]);
                        yE[0] = fma(a8[0], YE[c], w8[0]);
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            yA[i] = fma(a8[i], yA[i - 1], vr[8 * c + i]);
                            yE[i] = fma(a8[i], yE[i - 1], w8[i]);
                        }
                        tmem::wait_ld_dep(g8);
                        tmem::wait_ld_dep(d8);
                        double uA = UA[c], uE = UE[c];
#pragma unroll
                        for (int i2 = 3; i2 >= 0; --i2) {
                            const double2 pp = lds_v2f64(a_p + (c * 4 + i2) * 512);
                            {
                                const int i = 2 * i2 + 1;
                                uA = fma(g8[i], uA, yA[i]);
                                uE = fma(g8[i], uE, yE[i]);
                                const double rA = fma(d8[i], uA, -vr[8 * c + i]);
                                w8[i] = fma(d8[i], uE, -w8[i]);
                                vr[8 * c + i] = max_like_std(rA, pp.y);
                            }
                            {
                                const int i = 2 * i2;
                                uA = fma(g8[i], uA, yA[i]);
                                uE = fma(g8[i], uE, yE[i]);
                                const double rA = fma(d8[i], uA, -vr[8 * c + i]);
                                w8[i] = fma(d8[i], uE, -w8[i]);
                                vr[8 * c + i] = max_like_std(rA, pp.x);
                            }
                        }
                        tmem::st8(tbase + T_VE + 16 * c, w8);
                        // next step's local sweeps of both solutions
                        yA[0] = vr[8 * c];
                        yE[0] = w8[0];
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            yA[i] = fma(a8[i], yA[i - 1], vr[8 * c + i]);
                            yE[i] = fma(a8[i], yE[i - 1], w8[i]);
                        }
                        eA[c] = yA[7];
                        eE[c] = yE[7];
                        uA = yA[7];
                        uE = yE[7];
#pragma unroll
                        for (int i = 6; i >= 0; --i) {
                            uA = fma(g8[i], uA, yA[i]);
                            uE = fma(g8[i], uE, yE[i]);
                        }
                        fA[c] = uA;
                        fE[c] = uE;
                    }
                }
                tmem::wait_st();
            };
            switch (levels) {
                case 1: march(std::integral_constant<int, 1>{}); break;
                case 2: march(std::integral_constant<int, 2>{}); break;
                case 3: march(std::integral_constant<int, 3>{}); break;
                case 4: march(std::integral_constant<int, 4>{}); break;
                default: march(std::integral_constant<int, 5>{}); break;
            }
            if (lane == 0) {
                const int bucket = B.max_mode == 0 ? 0 : (levels == 5 ? 1 : 6 - levels);
                atomicAdd(&B.status[2 + bucket], 1u);
            }
            // ---------------- epilogue: both final solutions to shared memory, x_j recomputed ------
            double* vfinA = myp;            // the floors are not needed any more
            double* vfinE = st + warp * N;  // the stage is free: every set-up of the group is finished
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NODES; ++i) vfinA[lane * NODES + i] = vr[i];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                double w8[8];
                tmem::ld8(tbase + T_VE + 16 * c, w8);
                tmem::wait_ld_dep(w8);
#pragma unroll
                for (int i = 0; i < 8; ++i) vfinE[lane * NODES + 8 * c + i] = w8[i];
            }
            __syncwarp();
            {
                const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + my_pde) : my_pde;
                const PdeScalars sc = pde_scalars(load_option(B.opts + rep), B);
                uint32_t q0, q1;
                chain_range(B, my_pde, q0, q1);
                for (uint32_t q = q0 + lane; q < q1; q += 32) {
                    const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                    price_option2(
                        B, oi, [&](int j) { return x_node(sc, B.density, j); }, [&](int j) { return vfinA[j]; },
                        [&](int j) { return vfinE[j]; });
                }
            }
        }
        __syncthreads();
    }
    tmem::fence_before();
    __syncthreads();
    if (warp == 0) tmem::dealloc<64 * NCH>(s_taddr);
}

}  // namespace kwfd1d
