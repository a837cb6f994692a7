// fd1d_warpf.cuh -- Layout W with the march in a template type F: the fp32 march of
// FD1D.GPU.PRECISION = f32 (fp64 set-up).  Same kernel as fd1d_warp.cuh's chunk-pair
// form with `double` -> F in the march, 8 instead of 16 TMEM columns per 8-value block and 32-bit
// shuffles; kept as a separate template because any edit of the fp64 kernel's source perturbs the
// register allocation of its hot loop (DESIGN.md "Tried and dropped").
#pragma once
#include "fd1d_warp.cuh"

namespace kwfd1d {

// F: arithmetic type of the march (double, or float with the set-up still in fp64: FD1D.GPU.PRECISION = f32)
template <int NCH, int MINB, bool ICMP, bool PAIR, class F>
__global__ void __launch_bounds__(128, MINB) fd1d_warpf_kernel(const Fd1dBatch B)
{
    static_assert(std::is_same<F, double>::value || PAIR, "the fp32 march exists in the chunk-pair form only");
    constexpr bool F64 = std::is_same<F, double>::value;
    constexpr int CW = F64 ? 16 : 8;  // TMEM columns of an 8-value block
    static_assert(NCH == 4 || (NCH == 2 && PAIR), "4 chunks per lane (512 < x <= 1024) or 2 (256 < x <= 512)");
    using L = WarpSmem<NCH>;
    constexpr int N = L::N;
    constexpr int P = L::P;   // set-up threads per PDE (Layout B's cooperative set-up)
    constexpr int G = L::G;   // PDEs set up side by side by the CTA's 128 threads
    constexpr int M = 8;
    constexpr int NODES = 8 * NCH;  // per lane

    extern __shared__ double smem[];
    double* xs = smem;                // [4][N]
    double* st_all = xs + 4 * N;          // [G][5][N]: a, g, D, p, v of the PDEs being set up; later the final v
    double* stA_all = st_all + G * 5 * N; // [G][3][P]: chunk products of a~, of g~, response R0
    double* scr_all = stA_all + G * 3 * P;  // [G][8 * P]
    double* misc = scr_all + G * 8 * P;   // [16] spare
    double* wconst = misc + 16;       // [4 warps][22][32 lanes] scan constants of the warp's PDE

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sg = threadIdx.x / P;   // set-up group of this thread
    const int k = threadIdx.x % P;    // its chunk in the group's PDE
    const int xDim = B.xDim;
    const int nsteps = B.tDim - 1;
    double* st = st_all + sg * (5 * N);
    double* st_A = stA_all + sg * (3 * P);
    double* st_G = st_A + P;
    double* st_R0 = st_G + P;
    double* scr = scr_all + sg * (8 * P);

    // tensor memory: 4 arrays x 8*NCH doubles per lane = 64*NCH columns per warp
    __shared__ uint32_t s_taddr;
    if (warp == 0) tmem::alloc<4 * CW * NCH>(smem_addr(&s_taddr));
    tmem::fence_before();
    __syncthreads();
    tmem::fence_after();
    const uint32_t tbase = s_taddr + ((uint32_t)(warp & 3) << 21);
    constexpr uint32_t T_A = 0, T_G = CW * NCH, T_D = 2 * CW * NCH, T_P = 3 * CW * NCH;  // column offsets, CW per chunk

    const uint32_t n_pde = batch_n_pde(B);
    const uint32_t n_grp = (n_pde + 3) / 4;
    for (uint32_t grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        F vr[NODES];                   // this lane's nodes of PDE 4*grp + warp
        double Ac[NCH], Gc[NCH], R0c[NCH];
        double bmax_mine = 1.;
        bool put_mine = true;

        // ---------------- set-up: G PDEs at a time, P threads each --------------------------------
        for (int q0 = 0; q0 < 4; q0 += G) {
            if (4 * grp + q0 >= n_pde) break;  // uniform across the CTA
            const int q = q0 + sg;
            const bool q_valid = 4 * grp + q < n_pde;
            const uint32_t pde = q_valid ? 4 * grp + q : n_pde - 1;  // a missing PDE is set up as a copy (barriers stay uniform)
            const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + pde) : pde;
            const kw_option opt = load_option(B.opts + rep);
            const PdeScalars sc = pde_scalars(opt, B);
            {
                double v[M], pj[M], a[M], g[M], D[M];
                setup_lu<M, P>(B, sc, ICMP ? -0. : -CUDART_INF, xs + q * N, scr, v, pj, a, g, D);
                // chunk scalars: A = prod a~, G = prod g~, R0 = d(u~_first)/d(Yin) (backward sweep of the prefix products)
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
                if (lane == 0) scr[k >> 5] = bmax;
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
            if (warp >= q0 && warp < q0 + G && 4 * grp + warp < n_pde) {
                // the owner pulls its lane's NCH chunks: coefficient arrays into TMEM, v into registers;
                // warp q0 + j owns the PDE that set-up group j just prepared
                const double* st = st_all + (warp - q0) * (5 * N);
                const double* st_A = stA_all + (warp - q0) * (3 * P);
                const double* st_G = st_A + P;
                const double* st_R0 = st_G + P;
                const double* scr = scr_all + (warp - q0) * (8 * P);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int ch = lane * NCH + c;
                    F t8[8];
#pragma unroll
                    for (int arr = 0; arr < 4; ++arr) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) t8[i] = (F)st[arr * N + ch * 8 + i];
                        tmem::st8(tbase + CW * NCH * arr + CW * c, t8);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) vr[8 * c + i] = (F)st[4 * N + ch * 8 + i];
                    Ac[c] = st_A[ch];
                    Gc[c] = st_G[ch];
                    R0c[c] = st_R0[ch];
                }
                tmem::wait_st();
                double bm = scr[0];
#pragma unroll
                for (int w = 1; w < P / 32; ++w) bm = fmax(bm, scr[w]);
                bmax_mine = bm;
                // the owner's PDE may differ from the one this thread helped to set up: read its flag directly
                const uint32_t rep_own = B.pde_rep ? __ldg(B.pde_rep + 4 * grp + warp) : 4 * grp + warp;
                put_mine = load_option(B.opts + rep_own).w < 0;
            }
            __syncthreads();
        }

        const uint32_t my_pde = 4 * grp + warp;
        const bool have = my_pde < n_pde;  // warp-uniform
        int levels = 5;
        if (have) {
            // ---------------- cross-lane scan multipliers (lane aggregates) ----------------------
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
            // ---------------- how many levels carry anything (DESIGN.md "Truncation") ------------
            {
                const double tol = (F64 ? 0x1p-56 : 0x1p-30) / (bmax_mine * (double)B.tDim);
                const double* x = xs + warp * N;
                const double x_here = fmax(0., x[min(lane * NODES, xDim - 1)]);
                levels = 0;
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const int src = min((lane + (1 << d)) * NODES + NODES - 1, xDim - 1);
                    const double growth = put_mine ? 1. : exp(fmax(0., x[src]) - x_here);
                    const bool bad = !(fabs(AfL[d]) <= tol) || !(fabs(GbL[d]) * growth <= tol);
                    if (__any_sync(FULL, bad)) levels = d + 1;
                }
                if (B.max_mode <= 1) levels = 5;  // FD1D.GPU.EXACT >= 1: every level
                if (levels < 1) levels = 1;
            }

            // ---------------- time march: no barrier ---------------------------------------------
            // The 22 per-lane scan constants are parked in shared memory ([const][lane], conflict-free)
            // and re-read every step: registers are for v, one chunk's sweeps and the coefficient
            // stage (current chunk + the prefetched next one).
            double* wc = wconst + warp * (22 * 32) + lane;
            const uint32_t a_wc = smem_addr(wc);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                sts_t<F>(a_wc + (0 + c) * 256, (F)Ac[c]);
                sts_t<F>(a_wc + (4 + c) * 256, (F)Gc[c]);
                sts_t<F>(a_wc + (8 + c) * 256, (F)R0c[c]);
            }
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                sts_t<F>(a_wc + (12 + d) * 256, (F)AfL[d]);
                sts_t<F>(a_wc + (17 + d) * 256, (F)GbL[d]);
            }
            __syncwarp();
            auto K = [&](int idx) { return lds_t<F>(a_wc + idx * 256); };

            auto march = [&](auto lev_c) {
                constexpr int LEV = decltype(lev_c)::value;
                F e[NCH], f[NCH];
                // local sweeps of chunk c from zero: e = last forward value, f = first backward value
                auto local = [&](const F (&a8)[8], const F (&g8)[8], int c) {
                    F y[8];
                    y[0] = vr[8 * c];
#pragma unroll
                    for (int i = 1; i < 8; ++i) y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                    e[c] = y[7];
                    F u = y[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) u = fma(g8[i], u, y[i]);
                    f[c] = u;
                };
                F an[8];  // a~ of the chunk that is processed next, always one load ahead
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    F a8[8], g8[8];
                    tmem::ld8(tbase + T_A + CW * c, a8);
                    tmem::ld8(tbase + T_G + CW * c, g8);
                    tmem::wait_ld_dep(a8);
                    tmem::wait_ld_dep(g8);
                    local(a8, g8, c);
                }
                tmem::ld8(tbase + T_A, an);  // in flight across the scans
                for (int step = 0; step < nsteps; ++step) {
                    // ---- forward: lane aggregate, scan over lanes, chunk-entry values
                    F S = e[0];
#pragma unroll
                    for (int c = 1; c < NCH; ++c) S = fma(K(c), S, e[c]);
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const F o = __shfl_up_sync(FULL, S, 1 << d);
                        S = fma(K(12 + d), o, S);
                    }
                    F Yin[NCH];
                    {
                        const F o = __shfl_up_sync(FULL, S, 1);
                        Yin[0] = lane ? o : F(0);
                    }
#pragma unroll
                    for (int c = 1; c < NCH; ++c) Yin[c] = fma(K(c - 1), Yin[c - 1], e[c - 1]);
                    // ---- backward: chunk-start values with the true forward carry, scan, chunk-exit values
#pragma unroll
                    for (int c = 0; c < NCH; ++c) f[c] = fma(K(8 + c), Yin[c], f[c]);
                    F T = f[NCH - 1];
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) T = fma(K(4 + c), T, f[c]);
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const F o = __shfl_down_sync(FULL, T, 1 << d);
                        T = fma(K(17 + d), o, T);
                    }
                    F Uin[NCH];
                    {
                        const F o = __shfl_down_sync(FULL, T, 1);
                        Uin[NCH - 1] = lane < 31 ? o : F(0);
                    }
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) Uin[c] = fma(K(4 + c + 1), Uin[c + 1], f[c + 1]);
                    // ---- per chunk: true sweeps from (Yin, Uin), projection, next step's local sweeps;
                    //      the next chunk's coefficients are already in flight
                    if constexpr (PAIR) {
#pragma unroll
                        for (int h = 0; h < NCH; h += 2) {
                            const int cA = h, cB = h + 1;
                            F aA[8], aB[8], gA[8], gB[8], dA[8], dB[8], pA[8], pB[8];
                            KW_W_LD8(tbase + T_A + CW * cA, aA);
                            KW_W_LD8(tbase + T_A + CW * cB, aB);
                            tmem::hot_wait(aA);
                            tmem::hot_wait(aB);
                            KW_W_LD8(tbase + T_G + CW * cA, gA);
                            KW_W_LD8(tbase + T_G + CW * cB, gB);
                            KW_W_LD8(tbase + T_D + CW * cA, dA);
                            KW_W_LD8(tbase + T_D + CW * cB, dB);
                            KW_W_LD8(tbase + T_P + CW * cA, pA);
                            KW_W_LD8(tbase + T_P + CW * cB, pB);
                            F yA[8], yB[8];
                            yA[0] = fma(aA[0], Yin[cA], vr[8 * cA]);
                            yB[0] = fma(aB[0], Yin[cB], vr[8 * cB]);
#pragma unroll
                            for (int i = 1; i < 8; ++i) {
                                yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                                yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                            }
                            tmem::hot_wait(gA);
                            tmem::hot_wait(gB);
                            tmem::hot_wait(dA);
                            tmem::hot_wait(dB);
                            tmem::hot_wait(pA);
                            tmem::hot_wait(pB);
                            F uA = Uin[cA], uB = Uin[cB];
#pragma unroll
                            for (int i = 7; i >= 0; --i) {
                                uA = fma(gA[i], uA, yA[i]);
                                uB = fma(gB[i], uB, yB[i]);
                                const F rA = fma(dA[i], uA, -vr[8 * cA + i]);
                                const F rB = fma(dB[i], uB, -vr[8 * cB + i]);
                                vr[8 * cA + i] = ICMP ? max_like_icmp(rA, pA[i]) : max_like_std(rA, pA[i]);
                                vr[8 * cB + i] = ICMP ? max_like_icmp(rB, pB[i]) : max_like_std(rB, pB[i]);
                            }
                            // next step's local sweeps of both chunks, interleaved
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
                    } else {
                    // TMEM loads are issued one block ahead of their use: g~, D, p of this chunk and a~ of
                    // the next arrive while the forward sweep's dependent chain runs
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        F a8[8], g8[8], d8[8], p8[8];
                        tmem::wait_ld_dep(an);
#pragma unroll
                        for (int i = 0; i < 8; ++i) a8[i] = an[i];
                        tmem::ld8(tbase + T_G + CW * c, g8);
                        tmem::ld8(tbase + T_D + CW * c, d8);
                        tmem::ld8(tbase + T_P + CW * c, p8);
                        tmem::ld8(tbase + T_A + CW * ((c + 1) & (NCH - 1)), an);
                        F y[8];
                        y[0] = fma(a8[0], Yin[c], vr[8 * c]);
#pragma unroll
                        for (int i = 1; i < 8; ++i) y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                        tmem::wait_ld_dep(g8);
                        tmem::wait_ld_dep(d8);
                        tmem::wait_ld_dep(p8);
                        F u = Uin[c];
#pragma unroll
                        for (int i = 7; i >= 0; --i) {
                            u = fma(g8[i], u, y[i]);
                            const F r = fma(d8[i], u, -vr[8 * c + i]);
                            vr[8 * c + i] = ICMP ? max_like_icmp(r, p8[i]) : max_like_std(r, p8[i]);
                        }
                        local(a8, g8, c);
                    }
                    }
                }
                if constexpr (!PAIR) tmem::wait_ld_dep(an);  // nothing in flight when the arrays are rewritten
            };
            switch (levels) {
                case 1: march(std::integral_constant<int, 1>{}); break;
                case 2: march(std::integral_constant<int, 2>{}); break;
                case 3: march(std::integral_constant<int, 3>{}); break;
                case 4: march(std::integral_constant<int, 4>{}); break;
                default: march(std::integral_constant<int, 5>{}); break;
            }
            if (lane == 0) {
                // histogram buckets shared with Layout B: 0 = exact requested, 1 = all 5 levels, 2/3/4 = 4/3/2, 5 = 1 level
                const int bucket = B.max_mode == 0 ? 0 : (levels == 5 ? 1 : 6 - levels);
                atomicAdd(&B.status[2 + bucket], 1u);
            }
            // ---------------- epilogue: interpolate every option of this chain -------------------
            double* vfin = st_all + warp * N;  // the stage is free: set-up finished before the march
#pragma unroll
            for (int i = 0; i < NODES; ++i) vfin[lane * NODES + i] = (double)vr[i];
            __syncwarp();
            {
                const double* x = xs + warp * N;
                uint32_t q0, q1;
                chain_range(B, my_pde, q0, q1);
                for (uint32_t q = q0 + lane; q < q1; q += 32) {
                    const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                    price_option(B, oi, [&](int j) { return x[j]; }, [&](int j) { return vfin[j]; });
                }
            }
        }
        __syncthreads();  // stage and x grids are rewritten by the next group
    }
    tmem::fence_before();
    __syncthreads();
    if (warp == 0) tmem::dealloc<4 * CW * NCH>(s_taddr);
}

}  // namespace kwfd1d
