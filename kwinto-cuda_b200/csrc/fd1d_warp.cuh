// fd1d_warp.cuh -- Layout W: one WARP per PDE, every lane owns 8*NCH contiguous nodes, the
// time-invariant coefficient arrays live in tensor memory, the time march has no barrier.
//
// Same scheme and same algebra as Layout B (fd1d_reg.cuh; reference src/Math/kwFd1d.cpp:61-136 +
// src/Math/kwMath.cpp:16-49 with the hoisted constant-dt LU in pivot-scaled unknowns), re-mapped so
// that the FP64 pipe is fed from instruction-level parallelism instead of resident warps:
//   * a lane's nodes are NCH chunks of 8; the chunk sweeps of one lane are independent dependent
//     chains, so NCH of them interleave in one instruction stream;
//   * the chunk-boundary values Yin_c / Uin_c come from an in-lane recurrence over the NCH chunks
//     and ONE Kogge-Stone scan over the 32 lanes per direction -- the lanes are 8*NCH nodes apart,
//     so the precomputed multipliers decay 8*NCH nodes per lane and 2-3 shuffle levels carry
//     everything that is not provably below 2^-56 of the solution scale;
//   * with the true incoming values known, every chunk is swept again from them ("dot + true
//     sweep"): y_i = a_i y_{i-1} + v_i from Yin_c, u_i = g_i u_{i+1} + y_i from Uin_c,
//     v'_i = max(D_i u_i - v_i, p_i).  Per node-step: 2 + 2 sweep DFMAs, 1 combine, 1 compare --
//     the same 5 + 1 as Layout B's local sweeps + two fix-ups, but only a~, g~, D and p are needed,
//     and the scans cost 16 + 2L DFMAs per PDE-step instead of 4 x 13;
//   * a~, g~, D, p (4 x 8*NCH doubles per lane) sit in TMEM (tmem.cuh) and are streamed with
//     tcgen05.ld in 8-double pieces right before use; v stays in registers.  No shared memory and no
//     __syncthreads inside the march; warps of a CTA only meet for set-up and the epilogue.
// A CTA is 4 warps = 4 PDEs (one per TMEM lane quarter); set-up of each PDE is done by all 128
// threads with Layout B's code (setup_lu: Moebius-composed pivots in the reference's operation
// order) and handed to the owning warp through a shared-memory stage.
#pragma once
#include <type_traits>

#include "fd1d_reg.cuh"

// width of the TMEM loads in the chunk-pair phase (ld8 = one x16, ld8_by4 = two x8, ld8_by2 = four x4)
#ifndef KW_W2_ROT
#define KW_W2_ROT 1
#endif
#ifndef KW_W2_ROTC
#define KW_W2_ROTC 0
#endif
#ifndef KW_W_LD8
#define KW_W_LD8 tmem::ld8
#endif

namespace kwfd1d {

// Index swizzle of the set-up stage.  The set-up thread of chunk k writes nodes 8k .. 8k+7 and the owning
// warp's lane l reads nodes 32l .. 32l+31 (NCH = 4), i.e. both sides walk shared memory with a stride of 64 resp.
// 256 bytes between lanes: 8- and 32-way bank conflicts (the owner's pull was 20 % of the set-up time,
// profiles/r1_au_*).  XOR-ing the low four bits of the node index with bits of j >> 4 and j >> 5 makes both
// patterns conflict-free per half-warp; it permutes inside aligned groups of 16 doubles, so arrays stay N long.
__device__ __forceinline__ int stage_swz(int j) { return j ^ ((j >> 4) & 7) ^ ((j >> 5) & 15); }

template <int NCH, int CTA = 128>
struct WarpSmem {
    static constexpr int N = 8 * NCH * 32;  // nodes per PDE tile
    static constexpr int P = 32 * NCH;      // set-up threads per PDE
    static constexpr int G = CTA / P;       // PDEs set up side by side
    // doubles: xs[4][N] | stage a, g, D, p, v [G][5][N] | chunk scalars A, G, R0 [G][3][P] | scratch [G][8*P]
    //          | misc [16] | per-warp scan constants [4][22][32]
    static constexpr size_t bytes()
    {
        return sizeof(double) * (size_t)(4 * N + G * 5 * N + G * 3 * P + G * 8 * P + 16 + 4 * 22 * 32);
    }
};

// PAIR: the chunk phase processes two chunks in lock step (two interleaved dependent chains per warp)
// instead of one chunk at a time with its coefficients prefetched.
// BS: the fused march of the control-variate pricer "FD1D-BS" (Fd1d_BlackScholes_Pricer::price, reference
// src/Pricer/kwFd1d_BlackScholes.cpp:15-43: the solve as given plus the solve of a European copy of every
// chain).  Both solves share the grid and the hoisted LU, so ONE set-up and ONE tensor-memory copy of a~, g~,
// D serve both; the European march has no floor loads, compares or selects (308 instead of 444 instructions
// per step).  Prices go to B.prices (as given) and B.prices_eu (European); capi.cu adds the closed form.
//   BS = 2 (variant 253): every warp marches its PDE as given, then re-creates the payoff from the x grid
//           and marches the European copy with the coefficients still in tensor memory.  A chain that is
//           given as European is marched once and priced into both arrays.
//   BS = 1 (variant 252): EIGHT warps per CTA, warp w < 4 marches PDE w as given and warp w + 4 its European
//           copy at the same time (shared TMEM lane quarter w, two PDEs set up side by side, one CTA per SM).
//           Measured slower than BS = 2: the European warp finishes early and its partner runs on alone.
// RT: the number of scan levels is a run-time value inside ONE march loop instead of five specialised
// copies of it.  Warps of an SM then execute the same code whatever their PDE's level count: the SM's
// instruction cache holds ~32 KB, a specialised step is 7 KB, and with BS there are two roles as well
// (ncu stall_no_instruction 0.92 cycles per instruction with 6 hot loops per SM, profiles/r1_aw_*).
// SPLIT: every chunk-pair phase of a step is a basic block of its own (an always-true branch on a launch
// parameter ptxas cannot see through, Fd1dBatch::opq_lim) and a~, g~ are loaded a second time for the next step's
// local sweeps (through Fd1dBatch::opq_zero, so that the second tcgen05.ld is not merged with the first).  ptxas
// schedules inside basic blocks: without the split it hoists the tcgen05.ld of all four chunks to the top of the
// step, runs out of 16-register destination blocks and copies the survivors out (118 moves of 417 instructions);
// with it a value dies where it is used (8 moves of 342 instructions, no spills).
template <int NCH, int MINB, bool ICMP, bool PAIR = false, int BS = 0, bool RT = false, bool SPLIT = false>
__global__ void __launch_bounds__(BS == 1 ? 256 : 128, MINB) fd1d_warp_kernel(const Fd1dBatch B)
{
    static_assert(!SPLIT || PAIR, "SPLIT is a form of the chunk-pair phase");
    static_assert(NCH == 4 || (NCH == 2 && PAIR), "4 chunks per lane (512 < x <= 1024) or 2 (256 < x <= 512)");
    static_assert(!BS || (PAIR && !ICMP && RT), "fused FD1D-BS march: chunk pairs, run-time scan levels");
    static_assert(BS != 1 || NCH == 4, "BS = 1 (European copy in warp w + 4): 512 < x <= 1024 only");
    using L = WarpSmem<NCH, BS == 1 ? 256 : 128>;
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
    const int warp = BS == 1 ? (threadIdx.x >> 5) & 3 : threadIdx.x >> 5;  // PDE of the group = TMEM lane quarter
    const bool euro = BS == 1 && threadIdx.x >= 128;                       // BS = 1: the European copy's warp
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
    if (threadIdx.x < 32) tmem::alloc<64 * NCH>(smem_addr(&s_taddr));
    tmem::fence_before();
    __syncthreads();
    tmem::fence_after();
    const uint32_t tbase = s_taddr + ((uint32_t)(warp & 3) << 21);
    const uint32_t tbase2 = SPLIT ? tbase + B.opq_zero : tbase;  // the same address, opaque to the compiler
    constexpr uint32_t T_A = 0, T_G = 16 * NCH, T_D = 32 * NCH, T_P = 48 * NCH;  // column offsets, 16 per chunk

    const uint32_t n_pde = batch_n_pde(B);
    const uint32_t n_grp = (n_pde + 3) / 4;
    for (uint32_t grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        double vr[NODES];              // this lane's nodes of PDE 4*grp + warp
        double Ac[NCH], Gc[NCH], R0c[NCH];
        double bmax_mine = 1.;
        bool put_mine = true, amer_mine = true;

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
                    const int js = stage_swz(k * M + i);
                    st[0 * N + js] = a[i];
                    st[1 * N + js] = g[i];
                    st[2 * N + js] = D[i];
                    st[3 * N + js] = pj[i];
                    st[4 * N + js] = v[i];
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
                    double t8[8];
                    if (!euro) {
#pragma unroll
                        for (int arr = 0; arr < 4; ++arr) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) t8[i] = st[arr * N + stage_swz(ch * 8 + i)];
                            tmem::st8(tbase + 16 * NCH * arr + 16 * c, t8);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) vr[8 * c + i] = st[4 * N + stage_swz(ch * 8 + i)];
                    Ac[c] = st_A[ch];
                    Gc[c] = st_G[ch];
                    R0c[c] = st_R0[ch];
                }
                if (!euro) tmem::wait_st();
                double bm = scr[0];
#pragma unroll
                for (int w = 1; w < P / 32; ++w) bm = fmax(bm, scr[w]);
                bmax_mine = bm;
                // the owner's PDE may differ from the one this thread helped to set up: read its flag directly
                const uint32_t rep_own = B.pde_rep ? __ldg(B.pde_rep + 4 * grp + warp) : 4 * grp + warp;
                const kw_option own = load_option(B.opts + rep_own);
                put_mine = own.w < 0;
                amer_mine = own.e != 0;
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
                const double tol = 0x1p-56 / (bmax_mine * (double)B.tDim);
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
            if (!euro) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    wc[(0 + c) * 32] = Ac[c];
                    wc[(4 + c) * 32] = Gc[c];
                    wc[(8 + c) * 32] = R0c[c];
                }
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    wc[(12 + d) * 32] = AfL[d];
                    wc[(17 + d) * 32] = GbL[d];
                }
            }
            __syncwarp();
            if constexpr (BS == 1) {
                // the European warp reads its partner's tensor-memory arrays and scan constants
                tmem::fence_before();
                asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
                tmem::fence_after();
            }
            const uint32_t a_wc = smem_addr(wc);
            auto K = [&](int idx) { return lds_f64(a_wc + idx * 256); };

            auto march = [&](auto lev_c, auto euro_c) {
                constexpr int LEV = decltype(lev_c)::value;
                constexpr bool EURO = decltype(euro_c)::value;  // European copy: no floor, no compare
                double e[NCH], f[NCH];
                // local sweeps of chunk c from zero: e = last forward value, f = first backward value
                auto local = [&](const double (&a8)[8], const double (&g8)[8], int c) {
                    double y[8];
                    y[0] = vr[8 * c];
#pragma unroll
                    for (int i = 1; i < 8; ++i) y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                    e[c] = y[7];
                    double u = y[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) u = fma(g8[i], u, y[i]);
                    f[c] = u;
                };
                double an[8];  // a~ of the chunk that is processed next, always one load ahead
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    double a8[8], g8[8];
                    tmem::ld8(tbase + T_A + 16 * c, a8);
                    tmem::ld8(tbase + T_G + 16 * c, g8);
                    tmem::wait_ld_dep(a8);
                    tmem::wait_ld_dep(g8);
                    local(a8, g8, c);
                }
                tmem::ld8(tbase + T_A, an);  // in flight across the scans
                for (int step = 0; step < nsteps; ++step) {
                    // ---- forward: lane aggregate, scan over lanes, chunk-entry values
                    double S = e[0];
#pragma unroll
                    for (int c = 1; c < NCH; ++c) S = fma(K(c), S, e[c]);
                    if constexpr (LEV > 0) {
#pragma unroll
                        for (int d = 0; d < LEV; ++d) {
                            const double o = __shfl_up_sync(FULL, S, 1 << d);
                            S = fma(K(12 + d), o, S);
                        }
                    } else {
                        // run-time level count (warp-uniform, 1..5): nested so that the usual 1-3 levels skip the rest
                        auto up = [&](int d) {
                            const double o = __shfl_up_sync(FULL, S, 1 << d);
                            S = fma(K(12 + d), o, S);
                        };
                        up(0);
                        if (levels > 1) {
                            up(1);
                            if (levels > 2) {
                                up(2);
                                if (levels > 3) {
                                    up(3);
                                    if (levels > 4) up(4);
                                }
                            }
                        }
                    }
                    double Yin[NCH];
                    {
                        const double o = __shfl_up_sync(FULL, S, 1);
                        Yin[0] = lane ? o : 0.;
                    }
#pragma unroll
                    for (int c = 1; c < NCH; ++c) Yin[c] = fma(K(c - 1), Yin[c - 1], e[c - 1]);
                    // ---- backward: chunk-start values with the true forward carry, scan, chunk-exit values
#pragma unroll
                    for (int c = 0; c < NCH; ++c) f[c] = fma(K(8 + c), Yin[c], f[c]);
                    double T = f[NCH - 1];
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) T = fma(K(4 + c), T, f[c]);
                    if constexpr (LEV > 0) {
#pragma unroll
                        for (int d = 0; d < LEV; ++d) {
                            const double o = __shfl_down_sync(FULL, T, 1 << d);
                            T = fma(K(17 + d), o, T);
                        }
                    } else {
                        auto down = [&](int d) {
                            const double o = __shfl_down_sync(FULL, T, 1 << d);
                            T = fma(K(17 + d), o, T);
                        };
                        down(0);
                        if (levels > 1) {
                            down(1);
                            if (levels > 2) {
                                down(2);
                                if (levels > 3) {
                                    down(3);
                                    if (levels > 4) down(4);
                                }
                            }
                        }
                    }
                    double Uin[NCH];
                    {
                        const double o = __shfl_down_sync(FULL, T, 1);
                        Uin[NCH - 1] = lane < 31 ? o : 0.;
                    }
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) Uin[c] = fma(K(4 + c + 1), Uin[c + 1], f[c + 1]);
                    // ---- per chunk: true sweeps from (Yin, Uin), projection, next step's local sweeps;
                    //      the next chunk's coefficients are already in flight
                    if constexpr (PAIR) {
#pragma unroll
                        for (int h = 0; h < NCH; h += 2) {
                            if (SPLIT && !(step < B.opq_lim[h >> 1])) continue;  // never taken: basic-block boundary
                            const int cA = h, cB = h + 1;
                            double aA[8], aB[8], gA[8], gB[8], dA[8], dB[8], pA[8], pB[8];
                            KW_W_LD8(tbase + T_A + 16 * cA, aA);
                            KW_W_LD8(tbase + T_A + 16 * cB, aB);
                            tmem::hot_wait(aA, aB);
                            KW_W_LD8(tbase + T_G + 16 * cA, gA);
                            KW_W_LD8(tbase + T_G + 16 * cB, gB);
                            KW_W_LD8(tbase + T_D + 16 * cA, dA);
                            KW_W_LD8(tbase + T_D + 16 * cB, dB);
                            if constexpr (!EURO) {
                                KW_W_LD8(tbase + T_P + 16 * cA, pA);
                                KW_W_LD8(tbase + T_P + 16 * cB, pB);
                            }
                            double yA[8], yB[8];
                            yA[0] = fma(aA[0], Yin[cA], vr[8 * cA]);
                            yB[0] = fma(aB[0], Yin[cB], vr[8 * cB]);
#pragma unroll
                            for (int i = 1; i < 8; ++i) {
                                yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                                yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                            }
                            if constexpr (!EURO) {
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
                            // next step's local sweeps of both chunks, interleaved
                            if constexpr (SPLIT) {
                                KW_W_LD8(tbase2 + T_A + 16 * cA, aA);
                                KW_W_LD8(tbase2 + T_A + 16 * cB, aB);
                                KW_W_LD8(tbase2 + T_G + 16 * cA, gA);
                                KW_W_LD8(tbase2 + T_G + 16 * cB, gB);
                                tmem::hot_wait(aA, aB);
                                tmem::hot_wait(gA, gB);
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
                        continue;
                    }
                    // TMEM loads are issued one block ahead of their use: g~, D, p of this chunk and a~ of
                    // the next arrive while the forward sweep's dependent chain runs
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        double a8[8], g8[8], d8[8], p8[8];
                        tmem::wait_ld_dep(an);
#pragma unroll
                        for (int i = 0; i < 8; ++i) a8[i] = an[i];
                        tmem::ld8(tbase + T_G + 16 * c, g8);
                        tmem::ld8(tbase + T_D + 16 * c, d8);
                        tmem::ld8(tbase + T_P + 16 * c, p8);
                        tmem::ld8(tbase + T_A + 16 * ((c + 1) & (NCH - 1)), an);
                        double y[8];
                        y[0] = fma(a8[0], Yin[c], vr[8 * c]);
#pragma unroll
                        for (int i = 1; i < 8; ++i) y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                        tmem::wait_ld_dep(g8);
                        tmem::wait_ld_dep(d8);
                        tmem::wait_ld_dep(p8);
                        double u = Uin[c];
#pragma unroll
                        for (int i = 7; i >= 0; --i) {
                            u = fma(g8[i], u, y[i]);
                            const double r = fma(d8[i], u, -vr[8 * c + i]);
                            vr[8 * c + i] = ICMP ? max_like_icmp(r, p8[i]) : max_like_std(r, p8[i]);
                        }
                        local(a8, g8, c);
                    }
                }
                tmem::wait_ld_dep(an);  // nothing in flight when the arrays are rewritten
            };
            // BS = 2: the march as given keeps its level-specialised loops (7 % faster than the run-time form,
            // variant 235 vs 233); only the European march uses run-time levels -- at most 3 + 1 hot loops per SM
            if constexpr (RT && BS != 2) {
                if (euro)
                    march(std::integral_constant<int, 0>{}, std::integral_constant<bool, BS == 1>{});
                else
                    march(std::integral_constant<int, 0>{}, std::false_type{});
            } else {
                switch (levels) {
                    case 1: march(std::integral_constant<int, 1>{}, std::false_type{}); break;
                    case 2: march(std::integral_constant<int, 2>{}, std::false_type{}); break;
                    case 3: march(std::integral_constant<int, 3>{}, std::false_type{}); break;
                    case 4: march(std::integral_constant<int, 4>{}, std::false_type{}); break;
                    default: march(std::integral_constant<int, 5>{}, std::false_type{}); break;
                }
            }
            if (lane == 0 && !euro) {
                // histogram buckets shared with Layout B: 0 = exact requested, 1 = all 5 levels, 2/3/4 = 4/3/2, 5 = 1 level
                const int bucket = B.max_mode == 0 ? 0 : (levels == 5 ? 1 : 6 - levels);
                atomicAdd(&B.status[2 + bucket], 1u);
            }
            // ---------------- epilogue: interpolate every option of this chain -------------------
            double* vfin = st_all + (warp + (euro ? 4 : 0)) * N;  // the stage is free: set-up finished before the march
            const double* x = xs + warp * N;
            auto emit = [&](double* out) {
#pragma unroll
                for (int i = 0; i < NODES; ++i) vfin[lane * NODES + i] = vr[i];
                __syncwarp();
                Fd1dBatch Bo = B;
                Bo.prices = out;
                uint32_t q0, q1;
                chain_range(B, my_pde, q0, q1);
                for (uint32_t q = q0 + lane; q < q1; q += 32) {
                    const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                    price_option(Bo, oi, [&](int j) { return x[j]; }, [&](int j) { return vfin[j]; });
                }
                __syncwarp();  // vfin is rewritten by the next emit
            };
            emit(euro ? B.prices_eu : B.prices);
            if constexpr (BS == 2) {
                if (amer_mine) {
                    // the European copy: payoff again (src/Pricer/kwFd1d.cpp:127-139, as in setup_lu), same LU
                    // (still in tensor memory), no projection
#pragma unroll
                    for (int i = 0; i < NODES; ++i) {
                        const int j = lane * NODES + i;
                        vr[i] = j < xDim ? payoff_node(put_mine, x[j]) : 0.;
                    }
                    march(std::integral_constant<int, 0>{}, std::true_type{});
                }
                emit(B.prices_eu);  // a chain given as European: one march, both arrays
            }
        }
        if constexpr (BS == 1) tmem::fence_before();  // the partner's tcgen05.ld of this group are complete
        __syncthreads();  // stage, x grids and tensor memory are rewritten by the next group
        if constexpr (BS == 1) tmem::fence_after();
    }
    tmem::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem::dealloc<64 * NCH>(s_taddr);
}

// ---------------------------------------------------------------------------------------------
// Layout W, second arrangement: the solution v itself lives in tensor memory next to a~, g~, D
// (read for the sweeps, written back after the projection), and the projection floor p is read from
// shared memory two nodes at a time, just in time.  Registers then only hold what is in flight --
// two chunks' sweeps and coefficient blocks -- which leaves the instruction scheduler room to
// interleave the two dependent chains of a chunk pair instead of serialising them.
// Shared memory per CTA: one x grid (set-up only; the epilogue recomputes x_j with the same
// function), the set-up stage, one floor array per warp (re-used for the final v).
template <int NCH>
struct Warp2Smem {
    static constexpr int N = 8 * NCH * 32;
    static constexpr int P = 32 * NCH;
    // doubles: xs[N] | stage a, g, D, p, v [5][N] (setup scratch aliases the v slot) | floors [4][N]
    //          | chunk scalars A, G, R0 [3][P] | misc [16] | per-warp scan constants [4][22][32]
    static constexpr size_t bytes() { return sizeof(double) * (size_t)(N + 5 * N + 4 * N + 3 * P + 16 + 4 * 22 * 32); }
};

template <int NCH, int MINB, bool ICMP>
__global__ void __launch_bounds__(128, MINB) fd1d_warp2_kernel(const Fd1dBatch B)
{
    static_assert(NCH == 4, "4 chunks per lane (512 < x <= 1024)");
    using L = Warp2Smem<NCH>;
    constexpr int N = L::N;
    constexpr int P = L::P;
    constexpr int M = 8;
    constexpr int NODES = 8 * NCH;

    extern __shared__ double smem[];
    double* xs = smem;            // [N]
    double* st = xs + N;          // [5][N]
    double* scr = st + 4 * N;     // setup_lu scratch = the v slot of the stage (free until staging)
    double* ps = st + 5 * N;      // [4][N] floors, [chunk][node pair][lane] double2; after the march: final v
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
    constexpr uint32_t T_A = 0, T_G = 16 * NCH, T_D = 32 * NCH, T_V = 48 * NCH;
    // The v block is stored rotated by KW_W2_ROT doubles: a tcgen05.ld lands element j of every block in
    // the same register bank, and a DFMA whose two block operands (a~_i and v_i, D_i and v_i) share
    // a bank makes ptxas copy one of them elsewhere first (a third of the loop's instructions).
#define KW_RX(isv, c, i) (((i) + KW_W2_ROTC * ((c) & 1) + ((isv) ? KW_W2_ROT : 0)) & 7)

    const uint32_t n_pde = batch_n_pde(B);
    const uint32_t n_grp = (n_pde + 3) / 4;
    for (uint32_t grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        int levels = 5;
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
                setup_lu<M, P>(B, sc, ICMP ? -0. : -CUDART_INF, xs, scr, v, pj, a, g, D);
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
                // the owner pulls its lane's chunks: a~, g~, D, v into TMEM, the floor into its shared array
                double Ac[NCH], Gc[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int ch = lane * NCH + c;
                    double t8[8];
#pragma unroll
                    for (int arr = 0; arr < 4; ++arr) {
                        const int from = arr == 3 ? 4 : arr;  // TMEM slot 3 holds v (rotated)
#pragma unroll
                        for (int i = 0; i < 8; ++i) t8[KW_RX(arr == 3, c, i)] = st[from * N + ch * 8 + i];
                        tmem::st8(tbase + 16 * NCH * arr + 16 * c, t8);
                    }
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
                // cross-lane scan multipliers from the lane aggregates
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
                // how many levels carry anything (DESIGN.md "Truncation")
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
                double e[NCH], f[NCH];
                // first local sweeps
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    double a8[8], g8[8], v8[8];
                    tmem::ld8(tbase + T_A + 16 * c, a8);
                    tmem::ld8(tbase + T_G + 16 * c, g8);
                    tmem::ld8(tbase + T_V + 16 * c, v8);
                    tmem::wait_ld_dep(a8);
                    tmem::wait_ld_dep(g8);
                    tmem::wait_ld_dep(v8);
                    double y[8];
                    y[0] = v8[KW_RX(1, c, 0)];
#pragma unroll
                    for (int i = 1; i < 8; ++i) y[i] = fma(a8[KW_RX(0, c, i)], y[i - 1], v8[KW_RX(1, c, i)]);
                    e[c] = y[7];
                    double u = y[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) u = fma(g8[KW_RX(0, c, i)], u, y[i]);
                    f[c] = u;
                }
                for (int step = 0; step < nsteps; ++step) {
                    // ---- forward: lane aggregate, scan over lanes, chunk-entry values
                    double S = e[0];
#pragma unroll
                    for (int c = 1; c < NCH; ++c) S = fma(K(c), S, e[c]);
#pragma unroll
                    for (int d = 0; d < LEV; ++d) {
                        const double o = __shfl_up_sync(FULL, S, 1 << d);
                        S = fma(K(12 + d), o, S);
                    }
                    double Yin[NCH];
                    {
                        const double o = __shfl_up_sync(FULL, S, 1);
                        Yin[0] = lane ? o : 0.;
                    }
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
                    double Uin[NCH];
                    {
                        const double o = __shfl_down_sync(FULL, T, 1);
                        Uin[NCH - 1] = lane < 31 ? o : 0.;
                    }
#pragma unroll
                    for (int c = NCH - 2; c >= 0; --c) Uin[c] = fma(K(4 + c + 1), Uin[c + 1], f[c + 1]);
                    tmem::wait_st();  // last step's v is in place
                    // ---- chunk pairs: true sweeps from (Yin, Uin), projection, v back to TMEM, next local sweeps
#pragma unroll
                    for (int h = 0; h < NCH; h += 2) {
                        const int cA = h, cB = h + 1;
                        double aA[8], aB[8], vA[8], vB[8], gA[8], gB[8], dA[8], dB[8];
                        tmem::ld8(tbase + T_A + 16 * cA, aA);
                        tmem::ld8(tbase + T_A + 16 * cB, aB);
                        tmem::ld8(tbase + T_V + 16 * cA, vA);
                        tmem::ld8(tbase + T_V + 16 * cB, vB);
                        tmem::wait_ld_dep(aA);
                        tmem::wait_ld_dep(aB);
                        tmem::wait_ld_dep(vA);
                        tmem::wait_ld_dep(vB);
                        tmem::ld8(tbase + T_G + 16 * cA, gA);
                        tmem::ld8(tbase + T_G + 16 * cB, gB);
                        tmem::ld8(tbase + T_D + 16 * cA, dA);
                        tmem::ld8(tbase + T_D + 16 * cB, dB);
                        double yA[8], yB[8];
                        yA[0] = fma(aA[KW_RX(0, cA, 0)], Yin[cA], vA[KW_RX(1, cA, 0)]);
                        yB[0] = fma(aB[KW_RX(0, cB, 0)], Yin[cB], vB[KW_RX(1, cB, 0)]);
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            yA[i] = fma(aA[KW_RX(0, cA, i)], yA[i - 1], vA[KW_RX(1, cA, i)]);
                            yB[i] = fma(aB[KW_RX(0, cB, i)], yB[i - 1], vB[KW_RX(1, cB, i)]);
                        }
                        tmem::wait_ld_dep(gA);
                        tmem::wait_ld_dep(gB);
                        tmem::wait_ld_dep(dA);
                        tmem::wait_ld_dep(dB);
                        double uA = Uin[cA], uB = Uin[cB];
#pragma unroll
                        for (int i2 = 3; i2 >= 0; --i2) {
                            const double2 pA = lds_v2f64(a_p + (cA * 4 + i2) * 512);
                            const double2 pB = lds_v2f64(a_p + (cB * 4 + i2) * 512);
                            {
                                const int i = 2 * i2 + 1;
                                uA = fma(gA[KW_RX(0, cA, i)], uA, yA[i]);
                                uB = fma(gB[KW_RX(0, cB, i)], uB, yB[i]);
                                const double rA = fma(dA[KW_RX(0, cA, i)], uA, -vA[KW_RX(1, cA, i)]);
                                const double rB = fma(dB[KW_RX(0, cB, i)], uB, -vB[KW_RX(1, cB, i)]);
                                vA[KW_RX(1, cA, i)] = ICMP ? max_like_icmp(rA, pA.y) : max_like_std(rA, pA.y);
                                vB[KW_RX(1, cB, i)] = ICMP ? max_like_icmp(rB, pB.y) : max_like_std(rB, pB.y);
                            }
                            {
                                const int i = 2 * i2;
                                uA = fma(gA[KW_RX(0, cA, i)], uA, yA[i]);
                                uB = fma(gB[KW_RX(0, cB, i)], uB, yB[i]);
                                const double rA = fma(dA[KW_RX(0, cA, i)], uA, -vA[KW_RX(1, cA, i)]);
                                const double rB = fma(dB[KW_RX(0, cB, i)], uB, -vB[KW_RX(1, cB, i)]);
                                vA[KW_RX(1, cA, i)] = ICMP ? max_like_icmp(rA, pA.x) : max_like_std(rA, pA.x);
                                vB[KW_RX(1, cB, i)] = ICMP ? max_like_icmp(rB, pB.x) : max_like_std(rB, pB.x);
                            }
                        }
                        tmem::st8(tbase + T_V + 16 * cA, vA);
                        tmem::st8(tbase + T_V + 16 * cB, vB);
                        // next step's local sweeps of both chunks
                        yA[0] = vA[KW_RX(1, cA, 0)];
                        yB[0] = vB[KW_RX(1, cB, 0)];
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            yA[i] = fma(aA[KW_RX(0, cA, i)], yA[i - 1], vA[KW_RX(1, cA, i)]);
                            yB[i] = fma(aB[KW_RX(0, cB, i)], yB[i - 1], vB[KW_RX(1, cB, i)]);
                        }
                        e[cA] = yA[7];
                        e[cB] = yB[7];
                        uA = yA[7];
                        uB = yB[7];
#pragma unroll
                        for (int i = 6; i >= 0; --i) {
                            uA = fma(gA[KW_RX(0, cA, i)], uA, yA[i]);
                            uB = fma(gB[KW_RX(0, cB, i)], uB, yB[i]);
                        }
                        f[cA] = uA;
                        f[cB] = uB;
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
            // ---------------- epilogue: final v out of TMEM, x_j recomputed, interpolate -----------
            double* vfin = myp;  // the floors are not needed any more
            __syncwarp();
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                double v8[8];
                tmem::ld8(tbase + T_V + 16 * c, v8);
                tmem::wait_ld_dep(v8);
#pragma unroll
                for (int i = 0; i < 8; ++i) vfin[lane * NODES + 8 * c + i] = v8[KW_RX(1, c, i)];
            }
            __syncwarp();
            {
                const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + my_pde) : my_pde;
                const PdeScalars sc = pde_scalars(load_option(B.opts + rep), B);
                uint32_t q0, q1;
                chain_range(B, my_pde, q0, q1);
                for (uint32_t q = q0 + lane; q < q1; q += 32) {
                    const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                    price_option(B, oi, [&](int j) { return x_node(sc, B.density, j); }, [&](int j) { return vfin[j]; });
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
