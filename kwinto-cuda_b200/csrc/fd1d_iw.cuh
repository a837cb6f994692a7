// fd1d_iw.cuh -- Layout W with INDEPENDENT warps: every warp sets its own PDE up, marches it and prices
// its chain; the warps of a CTA share nothing but the tensor-memory allocation.
//
// Same scheme and algebra as fd1d_warp.cuh (reference src/Math/kwFd1d.cpp:61-136 + src/Math/kwMath.cpp:16-49,
// hoisted constant-dt LU in pivot-scaled unknowns, true-sweep formulation, chunk pairs, a~ g~ D p in tensor
// memory, v in registers) and the same arithmetic in the same order -- prices are bit-identical to
// fd1d_warp_kernel's.  What changes is who does the set-up and when:
//   * fd1d_warp_kernel sets the four PDEs of a CTA up one after the other with all 128 threads (Layout B's
//     setup_lu), five __syncthreads per PDE.  That phase is latency-bound (20 % issue utilisation) and while it
//     runs, only the SM's other CTA marches: measured fixed cost 2.0-2.7 ms of a 25.8 ms launch.
//   * here a lane owns its 8*NCH contiguous nodes from the first instruction on: x grid, payoff, rows of
//     B = 1 - dt/2 A, the Moebius-composed pivots (the lane's chunk maps -> one warp scan -> the reference's pivot
//     recurrence inside every chunk), a~, g~, D.  The rows are parked in the tensor-memory columns that will hold
//     a~, g~, D.  No shared-memory stage, no CTA barrier.  A warp in set-up issues little, so the sub-partition's
//     other warp marches at a lone warp's rate meanwhile; PDEs are handed out by an atomic counter, so there
//     is no wave tail either.
//   * the x grid is not kept: the scan-level test and the epilogue's interpolation recompute x_j with the
//     function the set-up used (same bits).
#pragma once
#include <type_traits>

#include "fd1d_reg.cuh"
#include "tmem.cuh"

namespace kwfd1d {

// out-of-line copies of the transcendental-heavy helpers: the set-up calls them from rolled loops
__device__ __noinline__ double x_node_ni(const PdeScalars& s, double density, int j) { return x_node(s, density, j); }
__device__ __noinline__ double payoff_node_ni(bool put, double x) { return payoff_node(put, x); }

template <int NCH>
struct IwSmem {
    static constexpr int N = 8 * NCH * 32;  // nodes per PDE tile
    static constexpr int SCR = 44 * 32;     // scratch per warp [row][lane]: half-chunk constants [5 * 4], carried half-chunk
                                            // values e_3, l_4 [2 * 4], chunk scalars of the set-up [12]; during the set-up the
                                            // pivots beta of the lane's 32 nodes (rows 0-27 and 40-43)
    // doubles per warp: final v of its PDE [N] (the payoff stage during set-up) | set-up scratch [SCR]
    static constexpr size_t bytes() { return sizeof(double) * (size_t)(4 * (N + SCR)); }
};

// BS: the fused march of the control-variate pricer "FD1D-BS" (Fd1d_BlackScholes_Pricer::price, reference
// src/Pricer/kwFd1d_BlackScholes.cpp:15-43: the solve as given plus the solve of a European copy of every chain).
// Both solves share the grid and the hoisted LU, so ONE set-up and ONE tensor-memory copy of a~, g~, D serve both:
// the warp marches its chain as given, prices into B.prices, re-creates the payoff and marches the European copy
// (no floor loads, compares or selects) into B.prices_eu.  A chain given as European is marched once and priced into
// both arrays.  capi.cu adds the closed form (bs_combine_kernel).
// D4: every 8-node sweep is split into two 4-node half-chains that run side by side.  The second half does not wait for
// the first: its entry value comes from a one- or two-DFMA lookahead through precomputed half-chunk products and the
// half-chunk results of the previous local sweeps,
//     forward   y_3* = e_3 + P_lo Yin              (e_3: lo half's local forward value, P_lo = a_0 a_1 a_2 a_3)
//     backward  u_4* = l_4 + R_hi y_3* + Q_hi Uin  (l_4: hi half's local backward value, R_hi, Q_hi = g_4 ... g_7)
// and the local sweeps are two 3-deep Horner halves combined by  e_c = e_hi + P_hi e_3,  f_c = l_lo + Q_lo (l_4 + R_hi e_3).
// 16 more DFMAs per step (186 instead of 170) for FOUR dependent chains per warp in every sweep phase instead of two and a
// dependent depth of 20 instead of 30 DFMAs per pair: the march is bound by the 8-cycle dependent-issue latency of the chains
// two warps can keep in flight, not by the FP64 pipe.  Five constants per chunk and the two carried
// half-chunk values live in the warp's shared scratch ([row][lane], one LDS / STS per use).
// PACK: PDEs per warp (1, 2 or 4).  A grid of at most 8*NCH*32 / PACK nodes takes only 32 / PACK lanes, so PACK PDEs ride
// in one warp side by side: lanes [h * LPP, (h + 1) * LPP) own PDE h of the warp's work unit.  The march needs no change at
// all: the first row of every PDE has no sub-diagonal (a~ = 0) and its last row no super-diagonal (g~ = 0), so the products the
// lane scans carry across a PDE boundary are exactly zero, and the shuffles are confined to a PDE's lanes by their width argument
// (nothing, not even a NaN, crosses over).  The step then costs what a full-tile PDE's step costs and advances PACK PDEs: the
// 512-node grid runs at the instruction mix of the 1024-node kernel (4 chunks per lane, two chunk pairs, rotated loop) instead of
// the single-pair loop of NCH = 2, whose scans have no sweeps to hide behind.
// F: arithmetic type of the march -- double, or float for FD1D.GPU.PRECISION = f32 (the set-up, the scan multipliers and the level
// test stay in fp64; a~, g~, D, the floor and v are rounded once, when the march takes them over; 8 instead of 16 tensor-memory
// columns per 8-value block, written over the parked fp64 rows chunk by chunk, always behind the reads).
template <int NCH, int MINB, bool BS = false, bool D4 = false, int PACK = 1, class F = double>
__global__ void __launch_bounds__(128, MINB) fd1d_iw_kernel(const Fd1dBatch B)
{
    static_assert(NCH == 4 || NCH == 2, "4 chunks per lane (512 < x <= 1024) or 2 (256 < x <= 512)");
    static_assert(PACK == 1 || PACK == 2 || PACK == 4, "1, 2 or 4 PDEs per warp");
    constexpr bool F64 = std::is_same<F, double>::value;
    static_assert(F64 || !D4, "the half-chunk lookahead exists in fp64 only");
    constexpr int CW = F64 ? 16 : 8;  // tensor-memory columns of an 8-value block of the march
    // fp32: a~, g~, D and the floor of the lane's 32 nodes are 128 registers -- with v (32) and the scan constants they fit the
    // register file, so the march loads them from tensor memory ONCE per PDE and its loop touches no memory at all; one basic
    // block per step, the compiler interleaves the sweeps of both chunk pairs (four dependent chains per warp)
    constexpr bool RC = !F64;
    constexpr int N = IwSmem<NCH>::N;
    constexpr int NODES = 8 * NCH;  // per lane
    constexpr int LPP = 32 / PACK;  // lanes per PDE
    constexpr int XT = N / PACK;    // nodes per PDE tile
    constexpr int MAXLEV = PACK == 1 ? 5 : (PACK == 2 ? 4 : 3);  // Kogge-Stone levels that stay inside a PDE

    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int pl = lane & (LPP - 1);  // lane within its PDE
    const int ph = lane / LPP;        // which PDE of the warp's work unit
    double* vfin = smem + warp * (N + IwSmem<NCH>::SCR);
    double* scr = vfin + N;
    const int xDim = B.xDim;
    const int nsteps = B.tDim - 1;

    // tensor memory: 4 arrays x 8*NCH doubles per lane = 64*NCH columns per warp
    __shared__ uint32_t s_taddr;
    if (threadIdx.x < 32) tmem::alloc<64 * NCH>(smem_addr(&s_taddr));
    tmem::fence_before();
    __syncthreads();
    tmem::fence_after();
    const uint32_t tbase = s_taddr + ((uint32_t)(warp & 3) << 21);
    const uint32_t tbase2 = tbase + B.opq_zero;  // the same address, opaque to the compiler (fd1d_warp.cuh, SPLIT)
    constexpr uint32_t T_A = 0, T_G = 16 * NCH, T_D = 32 * NCH, T_P = 48 * NCH;  // column offsets, 16 per chunk

    const uint32_t n_pde = batch_n_pde(B);
    for (;;) {
        uint32_t my_pde = 0;
        if (lane == 0) my_pde = atomicAdd(B.work_counter, 1u);  // work unit = PACK consecutive PDEs
        my_pde = __shfl_sync(FULL, my_pde, 0) * PACK;
        if (my_pde >= n_pde) break;
        my_pde += ph;
        const bool live = my_pde < n_pde;  // a short last unit: the spare lanes shadow the last PDE and emit nothing
        if (!live) my_pde = n_pde - 1;

        const uint32_t rep = B.pde_rep ? __ldg(B.pde_rep + my_pde) : my_pde;
        const kw_option opt = load_option(B.opts + rep);
        const PdeScalars sc = pde_scalars(opt, B);

        F vr[NODES];
        double Ac[NCH], Gc[NCH], R0c[NCH];
        double bmax = 0.;
        // ================= set-up (setup_lu of fd1d_reg.cuh, one warp, NCH chunks per lane) ==============
        // Rolled over the lane's chunks (the code runs once per PDE next to seven marching warps: a fully unrolled
        // set-up is 13 k instructions and evicts their hot loops from the instruction cache).  What crosses a chunk
        // boundary travels in a loop-carried scalar, in tensor memory (the rows of B, then 1/beta in the diagonal's
        // columns) or in the warp's shared scratch ([index][lane], conflict-free).
        {
            const int j0 = pl * NODES;
            double* s_v = vfin;            // [NODES][32] payoff, until the registers take it
            double* s_k = scr + 28 * 32;   // [3 * NCH][32] chunk scalars
            double* s_h = scr;             // [5 * NCH][32] half-chunk constants P_lo, P_hi, Q_lo, Q_hi, R_hi (D4)
            double* s_b = scr;             // [NODES][32] the pivots beta, until 1/beta is in tensor memory (rows 0-27, 40-43)
            auto beta_row = [](int n) { return n < 28 ? n : n + 12; };
            // ---- grid, payoff, projection floor, rows of B (parked in tensor memory)
            double bu_carry = 0.;
            {
                double x_m1 = pl ? x_node_ni(sc, B.density, j0 - 1) : 0.;
                double x_0 = x_node_ni(sc, B.density, j0);
                if (pl == 0) x_m1 = x_0;
                double inv_d = pl ? 1. / (x_0 - x_m1) : 0.;  // 1 / (x_j - x_{j-1}), handed from row to row (b_row_chained)
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    double xl[10];  // nodes 8c - 1 .. 8c + 8 of this lane
                    xl[0] = x_m1;
                    xl[1] = x_0;
#pragma unroll
                    for (int i = 1; i <= 8; ++i) xl[1 + i] = x_node_ni(sc, B.density, j0 + 8 * c + i);
                    if (pl == LPP - 1 && c == NCH - 1) xl[9] = xl[8];  // past the tile: the last node again
                    F t8[8];
                    double bl[8], bb[8], bu[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int j = j0 + 8 * c + i;
                        double p = 0.;
                        if (j < xDim) p = payoff_node_ni(sc.put, xl[1 + i]);
                        s_v[(8 * c + i) * 32 + lane] = p;
                        // projection skips the last node (src/Math/kwFd1d.cpp:130); European: never
                        t8[i] = (sc.american && j < xDim - 1) ? (F)p : (F)-CUDART_INF;
                        b_row_chained(sc, j, xDim, xl[i], xl[1 + i], xl[2 + i], inv_d, inv_d, bl[i], bb[i], bu[i]);
                    }
                    tmem::st8(tbase + T_P + CW * c, t8);
                    tmem::st8(tbase + T_A + 16 * c, bl);
                    tmem::st8(tbase + T_G + 16 * c, bb);
                    tmem::st8(tbase + T_D + 16 * c, bu);
                    bu_carry = bu[7];
                    x_m1 = xl[8];
                    x_0 = xl[9];
                }
                tmem::wait_st();
            }
            double bu_prev_lane = __shfl_up_sync(FULL, bu_carry, 1, LPP);
            if (pl == 0) bu_prev_lane = 0.;
            // ---- pivots: beta_j = b_j - c_j / beta_{j-1}, c_j = bl_j bu_{j-1}, as the Moebius map
            //      [[b_j, -c_j], [1, 0]] on (num; den); compose per chunk, per lane, scan over the lanes
            {
                Mat2 Lm = {1., 0., 0., 1.};
                bu_carry = bu_prev_lane;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    double bl[8], bb[8], bu[8];
                    tmem::ld8(tbase + T_A + 16 * c, bl);
                    tmem::ld8(tbase + T_G + 16 * c, bb);
                    tmem::ld8(tbase + T_D + 16 * c, bu);
                    tmem::wait_ld_dep(bl, bb, bu);
                    Mat2 m = {1., 0., 0., 1.};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const double cc = bl[i] * (i ? bu[i > 0 ? i - 1 : 0] : bu_carry);
                        Mat2 n;
                        n.m00 = fma(bb[i], m.m00, -cc * m.m10);
                        n.m01 = fma(bb[i], m.m01, -cc * m.m11);
                        n.m10 = m.m00;
                        n.m11 = m.m01;
                        m = n;
                        if ((i & 3) == 3) mat_normalise(m);
                    }
                    bu_carry = bu[7];
                    Lm = mat_mul(m, Lm);
                    mat_normalise(Lm);
                }
#pragma unroll
                for (int d = 1; d < LPP; d <<= 1) {
                    const Mat2 o = mat_shfl_up(Lm, d, LPP);
                    if (pl >= d) {
                        Lm = mat_mul(Lm, o);
                        mat_normalise(Lm);
                    }
                }
                const Mat2 E = mat_shfl_up(Lm, 1, LPP);
                double num = pl ? E.m00 : 1.;
                double den = pl ? E.m10 : 0.;
                // ---- The scan's pivots are a first guess only.  The recurrence beta_j = b_j - c_j / beta_{j-1} contracts
                //      (d beta_j / d beta_{j-1} = c_j / beta_{j-1}^2 < 1), so run serially it forgets its rounding errors,
                //      while the composed maps accumulate theirs over the whole grid (1e-13 relative at 4096 nodes: on stiff
                //      grids -- few time steps, dt/dx^2 in the thousands -- that alone moved prices by 3e-9 against the
                //      reference).  Polish: every lane re-runs the reference's recurrence over its own nodes from its
                //      incoming pivot and hands the result to the next lane, until no incoming pivot changes any more.  Lane l
                //      is exact after l sweeps at the latest; with the contraction, two or three sweeps do (stiff: up to ~10).
                //      The fixed point IS the serial recurrence of src/Math/kwMath.cpp:30-38, bit for bit.
                double pin = den == 0. ? CUDART_INF : num / den;  // pivot just before the lane's first node
#pragma unroll 1
                for (int sweep = 0; sweep < 34; ++sweep) {
                    double prev = pin;
                    bu_carry = bu_prev_lane;
#pragma unroll 1
                    for (int c = 0; c < NCH; ++c) {
                        double bl[8], bb[8], bu[8];
                        tmem::ld8(tbase + T_A + 16 * c, bl);
                        tmem::ld8(tbase + T_G + 16 * c, bb);
                        tmem::ld8(tbase + T_D + 16 * c, bu);
                        tmem::wait_ld_dep(bl, bb, bu);
                        const double was = s_b[beta_row(8 * c + 7) * 32 + lane];  // the previous sweep's pivot at the chunk's end
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            // the reference's order (src/Math/kwMath.cpp:32-33): gam = au[j-1] / bet;  bet = a[j] - al[j] * gam
                            const double gam = (i ? bu[i > 0 ? i - 1 : 0] : bu_carry) / prev;
                            prev = __dsub_rn(bb[i], __dmul_rn(bl[i], gam));
                            s_b[beta_row(8 * c + i) * 32 + lane] = prev;  // kept: the last sweep's pivots ARE the pivots
                        }
                        bu_carry = bu[7];
                        // From a chunk end where every lane reproduces the previous sweep's pivot bit for bit, the rest of the
                        // sweep would reproduce the rest of the previous one (same recurrence, same input): skip it.  The
                        // corrections of a sweep die out within a few nodes, so every sweep but the first stops after one chunk.
                        if (sweep > 0 && c < NCH - 1 &&
                            __all_sync(FULL, __double_as_longlong(prev) == __double_as_longlong(was))) {
                            prev = s_b[beta_row(NODES - 1) * 32 + lane];
                            break;
                        }
                    }
                    double pnew = __shfl_up_sync(FULL, prev, 1, LPP);
                    if (pl == 0) pnew = CUDART_INF;
                    const bool changed = __double_as_longlong(pnew) != __double_as_longlong(pin);
                    pin = pnew;
                    if (!__any_sync(FULL, changed)) break;
                }
                // ---- The sweep that found no incoming pivot changed ran the reference's recurrence from the final incoming pivots:
                //      its beta are the pivots.  1/beta replaces the diagonal in tensor memory (independent divisions, no chain);
                //      max |beta| scales the scan-truncation tolerance.
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    double ib[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const double beta = s_b[beta_row(8 * c + i) * 32 + lane];
                        ib[i] = 1. / beta;
                        bmax = fmax(bmax, fabs(beta));
                    }
                    tmem::st8(tbase + T_G + 16 * c, ib);
                }
                tmem::wait_st();
            }
            // ---- a~, g~, D into tensor memory (over the rows), chunk scalars
            {
                double ib[8];
                tmem::ld8(tbase + T_G, ib);
                tmem::wait_ld_dep(ib);
                double ib_last_lane, ib_first_lane = ib[0];
                {
                    double t[8];
                    tmem::ld8(tbase + T_G + 16 * (NCH - 1), t);
                    tmem::wait_ld_dep(t);
                    ib_last_lane = t[7];
                }
                double ib_prev = __shfl_up_sync(FULL, ib_last_lane, 1, LPP);    // 1/beta of the node before the chunk
                double ib_next_lane = __shfl_down_sync(FULL, ib_first_lane, 1, LPP);
                if (pl == 0) ib_prev = 0.;
                if (pl == LPP - 1) ib_next_lane = 0.;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    double bl[8], bu[8], ibn[8];
                    tmem::ld8(tbase + T_A + 16 * c, bl);
                    tmem::ld8(tbase + T_D + 16 * c, bu);
                    tmem::ld8(tbase + T_G + 16 * (c < NCH - 1 ? c + 1 : c), ibn);  // the next chunk's 1/beta
                    tmem::wait_ld_dep(bl, bu, ibn);
                    const double ib_next = c < NCH - 1 ? ibn[0] : ib_next_lane;
                    double a[8], g[8], D[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int j = j0 + 8 * c + i;
                        a[i] = -bl[i] * (i ? ib[i > 0 ? i - 1 : 0] : ib_prev);
                        g[i] = -bu[i] * (i < 7 ? ib[i < 7 ? i + 1 : i] : ib_next);
                        D[i] = j < xDim ? 2. * ib[i] : 0.;
                    }
                    if constexpr (F64) {
                        tmem::st8(tbase + T_A + 16 * c, a);
                        tmem::st8(tbase + T_G + 16 * c, g);
                        tmem::st8(tbase + T_D + 16 * c, D);
                    } else {
                        // 8 columns per block, over the first half of the fp64 rows that iteration c / 2 has consumed
                        // (8 c + 8 <= 16 c for c >= 1; c = 0: its own rows, already in registers); the next chunk's 1/beta
                        // at 16 (c + 1) is still ahead of the write
                        F af[8], gf[8], Df[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            af[i] = (F)a[i];
                            gf[i] = (F)g[i];
                            Df[i] = (F)D[i];
                        }
                        tmem::st8(tbase + T_A + CW * c, af);
                        tmem::st8(tbase + T_G + CW * c, gf);
                        tmem::st8(tbase + T_D + CW * c, Df);
                    }
                    // chunk scalars: A = prod a~, G = prod g~, R0 = d(u~_first)/d(Yin) (backward sweep of the prefix products)
                    double Pp[8];
                    Pp[0] = a[0];
#pragma unroll
                    for (int i = 1; i < 8; ++i) Pp[i] = a[i] * Pp[i - 1];
                    double Q0 = g[7], R0 = Pp[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) {
                        Q0 = g[i] * Q0;
                        R0 = fma(g[i], R0, Pp[i]);
                    }
                    s_k[(3 * c + 0) * 32 + lane] = Pp[7];
                    s_k[(3 * c + 1) * 32 + lane] = Q0;
                    s_k[(3 * c + 2) * 32 + lane] = R0;
                    if constexpr (D4) {
                        const double Ph4 = a[4], Ph5 = a[5] * Ph4, Ph6 = a[6] * Ph5, Ph7 = a[7] * Ph6;
                        double Rh = Ph7;  // response of the hi half's first backward value to its incoming forward value
                        Rh = fma(g[6], Rh, Ph6);
                        Rh = fma(g[5], Rh, Ph5);
                        Rh = fma(g[4], Rh, Ph4);
                        s_h[(5 * c + 0) * 32 + lane] = Pp[3];                       // P_lo = a_0 a_1 a_2 a_3
                        s_h[(5 * c + 1) * 32 + lane] = Ph7;                         // P_hi = a_4 a_5 a_6 a_7
                        s_h[(5 * c + 2) * 32 + lane] = ((g[0] * g[1]) * g[2]) * g[3];  // Q_lo
                        s_h[(5 * c + 3) * 32 + lane] = ((g[7] * g[6]) * g[5]) * g[4];  // Q_hi
                        s_h[(5 * c + 4) * 32 + lane] = Rh;
                    }
                    ib_prev = ib[7];
#pragma unroll
                    for (int i = 0; i < 8; ++i) ib[i] = ibn[i];
                }
                tmem::wait_st();
            }
#pragma unroll
            for (int d = LPP / 2; d >= 1; d >>= 1) bmax = fmax(bmax, __shfl_xor_sync(FULL, bmax, d));  // over the PDE's lanes
            __syncwarp();
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                Ac[c] = s_k[(3 * c + 0) * 32 + lane];
                Gc[c] = s_k[(3 * c + 1) * 32 + lane];
                R0c[c] = s_k[(3 * c + 2) * 32 + lane];
            }
#pragma unroll
            for (int i = 0; i < NODES; ++i) vr[i] = (F)s_v[i * 32 + lane];
            __syncwarp();  // the scratch is the final-v stage later
        }

        // ================= cross-lane scan multipliers (lane aggregates) ================================
        double AfL[5], GbL[5];
        {
            double AL = Ac[0], GL = Gc[0];
#pragma unroll
            for (int c = 1; c < NCH; ++c) {
                AL *= Ac[c];
                GL *= Gc[c];
            }
            double A = AL;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int s = 1 << d;
                const double o = __shfl_up_sync(FULL, A, s, LPP);
                AfL[d] = pl >= s ? A : 0.;
                if (pl >= s) A *= o;
            }
            double G = GL;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int s = 1 << d;
                const double o = __shfl_down_sync(FULL, G, s, LPP);
                GbL[d] = pl < LPP - s ? G : 0.;
                if (pl < LPP - s) G *= o;
            }
        }
        F AcF[NCH], GcF[NCH], R0F[NCH], AfF[5], GbF[5];  // the march's copies (fp64: the same registers), filled below
        auto kA = [&](int c) { return AcF[c]; };
        auto kG = [&](int c) { return GcF[c]; };
        auto kR = [&](int c) { return R0F[c]; };
        auto kAf = [&](int d) { return AfF[d]; };
        auto kGb = [&](int d) { return GbF[d]; };
        // ================= how many levels carry anything (DESIGN.md "Truncation") =======================
        // (PACK > 1: every PDE of the warp gets the level count it would get alone -- `own` -- and exact zeros as multipliers of
        // the levels past it, so that the warp's common count `levels` changes nothing for it: a PDE's prices do not depend on
        // which PDEs share its warp, bit for bit)
        int levels = 0, own = 0;
        {
            const double tol = (F64 ? 0x1p-56 : 0x1p-30) / (bmax * (double)B.tDim);
            const double x_here = fmax(0., x_node(sc, B.density, min(pl * NODES, xDim - 1)));
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int src = min((pl + (1 << d)) * NODES + NODES - 1, xDim - 1);
                const double growth = sc.put ? 1. : exp(fmax(0., x_node(sc, B.density, src)) - x_here);
                const bool bad = !(fabs(AfL[d]) <= tol) || !(fabs(GbL[d]) * growth <= tol);
                const unsigned votes = __ballot_sync(FULL, bad);
                if (votes) levels = d + 1;
                if ((votes >> (ph * LPP)) & (0xffffffffu >> (32 - LPP))) own = d + 1;
            }
            if (B.max_mode <= 1) levels = own = 5;  // FD1D.GPU.EXACT >= 1: every level
            if (levels > MAXLEV) levels = MAXLEV;  // (levels past a PDE's lanes carry exact zeros)
            if (levels < 1) levels = 1;
            if (own > MAXLEV) own = MAXLEV;
            if (own < 1) own = 1;
            if constexpr (PACK > 1) {
#pragma unroll
                for (int d = 0; d < 5; ++d)
                    if (d >= own) AfL[d] = GbL[d] = 0.;
            }
        }
        // keep the multipliers as values: the compiler would otherwise re-derive the first levels from Ac[] / Gc[]
        // inside the march loop (6 DMUL + the lane predicate per step) to save two registers
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            AcF[c] = (F)Ac[c];
            GcF[c] = (F)Gc[c];
            R0F[c] = (F)R0c[c];
        }
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            AfF[d] = (F)AfL[d];
            GbF[d] = (F)GbL[d];
            if constexpr (F64) asm volatile("" : "+d"(AfF[d]), "+d"(GbF[d]));
            else asm volatile("" : "+f"(AfF[d]), "+f"(GbF[d]));
        }

        // ================= time march: fd1d_warp_kernel's SPLIT chunk-pair form ==========================
        auto march = [&](auto lev_c, auto euro_c) {
            constexpr int LEV = decltype(lev_c)::value;
            constexpr bool EURO = decltype(euro_c)::value;  // European copy: no floor, no compare
            F e[NCH], f[NCH];
            // D4: half-chunk constants and carried half-chunk values in the warp's scratch, [row][lane]
            const uint32_t a_h = smem_addr(scr + lane);
            auto hP_lo = [&](int c) { return lds_f64(a_h + (5 * c + 0) * 256); };
            auto hP_hi = [&](int c) { return lds_f64(a_h + (5 * c + 1) * 256); };
            auto hQ_lo = [&](int c) { return lds_f64(a_h + (5 * c + 2) * 256); };
            auto hQ_hi = [&](int c) { return lds_f64(a_h + (5 * c + 3) * 256); };
            auto hR_hi = [&](int c) { return lds_f64(a_h + (5 * c + 4) * 256); };
            auto ld_e3 = [&](int c) { return lds_f64(a_h + (20 + c) * 256); };
            auto ld_l4 = [&](int c) { return lds_f64(a_h + (24 + c) * 256); };
            auto st_e3 = [&](int c, double v) { sts_f64(a_h + (20 + c) * 256, v); };
            auto st_l4 = [&](int c, double v) { sts_f64(a_h + (24 + c) * 256, v); };
            // local sweeps of chunk c from zero (next step's aggregates): e = last forward value, f = first backward value
            auto local_chunk = [&](int c, const F (&a8)[8], const F (&g8)[8]) {
                if constexpr (!D4) {
                    F y[8];
                    y[0] = vr[8 * c];
#pragma unroll
                    for (int i = 1; i < 8; ++i) y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                    e[c] = y[7];
                    F u = y[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) u = fma(g8[i], u, y[i]);
                    f[c] = u;
                } else {
                    double y[8];
                    y[0] = vr[8 * c];
                    y[4] = vr[8 * c + 4];
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        y[i] = fma(a8[i], y[i - 1], vr[8 * c + i]);
                        y[4 + i] = fma(a8[4 + i], y[3 + i], vr[8 * c + 4 + i]);
                    }
                    const double e3 = y[3];
                    e[c] = fma(hP_hi(c), e3, y[7]);
                    double ul = y[3], uh = y[7];
#pragma unroll
                    for (int i = 2; i >= 0; --i) {
                        ul = fma(g8[i], ul, y[i]);
                        uh = fma(g8[4 + i], uh, y[4 + i]);
                    }
                    f[c] = fma(hQ_lo(c), fma(hR_hi(c), e3, uh), ul);
                    st_e3(c, e3);
                    st_l4(c, uh);
                }
            };
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                F a8[8], g8[8];
                tmem::ld8(tbase + T_A + CW * c, a8);
                tmem::ld8(tbase + T_G + CW * c, g8);
                tmem::wait_ld_dep(a8, g8);
                local_chunk(c, a8, g8);
            }
            F aR[RC ? NODES : 1], gR[RC ? NODES : 1], dR[RC ? NODES : 1], pR[RC ? NODES : 1];
            if constexpr (RC) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    F t0[8], t1[8], t2[8], t3[8];
                    tmem::ld8(tbase + T_A + CW * c, t0);
                    tmem::ld8(tbase + T_G + CW * c, t1);
                    tmem::ld8(tbase + T_D + CW * c, t2);
                    tmem::ld8(tbase + T_P + CW * c, t3);
                    tmem::wait_ld_dep(t0, t1);
                    tmem::wait_ld_dep(t2, t3);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        aR[8 * c + i] = t0[i];
                        gR[8 * c + i] = t1[i];
                        dR[8 * c + i] = t2[i];
                        pR[8 * c + i] = t3[i];
                    }
                }
            }
            F Yin[NCH], Uin[NCH];
            // ---- the scans of one step: lane aggregates, Kogge-Stone over the lanes, chunk-entry / -exit values
            auto scan = [&]() {
                F S = e[0];
#pragma unroll
                for (int c = 1; c < NCH; ++c) S = fma(kA(c), S, e[c]);
#pragma unroll
                for (int d = 0; d < LEV; ++d) {
                    const F o = __shfl_up_sync(FULL, S, 1 << d, LPP);
                    S = fma(kAf(d), o, S);
                }
                // A PDE's first lane has no predecessor: __shfl_up hands it its own (finite) S back.  No select is needed -- whatever
                // Yin[0] is there, it is only ever multiplied by a~ of the grid's first node, A_0 or R0_0 of lane 0's first
                // chunk, all exactly 0 (a~_0 = -bl_0 / beta_{-1} with 1/beta_{-1} = 0).  Likewise Uin of lane 31's last chunk
                // below (g~ of the last node is 0).
                Yin[0] = __shfl_up_sync(FULL, S, 1, LPP);
#pragma unroll
                for (int c = 1; c < NCH; ++c) Yin[c] = fma(kA(c - 1), Yin[c - 1], e[c - 1]);
                // backward: chunk-start values with the true forward carry, scan, chunk-exit values
#pragma unroll
                for (int c = 0; c < NCH; ++c) f[c] = fma(kR(c), Yin[c], f[c]);
                F T = f[NCH - 1];
#pragma unroll
                for (int c = NCH - 2; c >= 0; --c) T = fma(kG(c), T, f[c]);
#pragma unroll
                for (int d = 0; d < LEV; ++d) {
                    const F o = __shfl_down_sync(FULL, T, 1 << d, LPP);
                    T = fma(kGb(d), o, T);
                }
                Uin[NCH - 1] = __shfl_down_sync(FULL, T, 1, LPP);
#pragma unroll
                for (int c = NCH - 2; c >= 0; --c) Uin[c] = fma(kG(c + 1), Uin[c + 1], f[c + 1]);
            };
            // ---- true forward sweeps of the chunk pair (cA, cA + 1) from Yin
            // (D4: y3[0], y3[1] = the lookahead values y_3* of the two chunks, needed again by the backward lookahead)
            auto fwd_pair = [&](int cA, F (&yA)[8], F (&yB)[8], F (&y3)[2]) {
                const int cB = cA + 1;
                F aA[8], aB[8];
                if constexpr (RC) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        aA[i] = aR[8 * cA + i];
                        aB[i] = aR[8 * cB + i];
                    }
                } else {
                    tmem::ld16(tbase + T_A + CW * cA, aA, aB);
                    tmem::hot_wait(aA, aB);
                }
                if constexpr (!D4) {
                    yA[0] = fma(aA[0], Yin[cA], vr[8 * cA]);
                    yB[0] = fma(aB[0], Yin[cB], vr[8 * cB]);
#pragma unroll
                    for (int i = 1; i < 8; ++i) {
                        yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                        yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                    }
                } else {
                    y3[0] = fma(hP_lo(cA), Yin[cA], ld_e3(cA));
                    y3[1] = fma(hP_lo(cB), Yin[cB], ld_e3(cB));
                    yA[0] = fma(aA[0], Yin[cA], vr[8 * cA]);
                    yB[0] = fma(aB[0], Yin[cB], vr[8 * cB]);
                    yA[4] = fma(aA[4], y3[0], vr[8 * cA + 4]);
                    yB[4] = fma(aB[4], y3[1], vr[8 * cB + 4]);
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                        yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                        yA[4 + i] = fma(aA[4 + i], yA[3 + i], vr[8 * cA + 4 + i]);
                        yB[4 + i] = fma(aB[4 + i], yB[3 + i], vr[8 * cB + 4 + i]);
                    }
                }
            };
            // ---- true backward sweeps from Uin, projection, next step's local sweeps (a~, g~ loaded again)
            auto back_pair = [&](int cA, F (&yA)[8], F (&yB)[8], const F (&y3)[2]) {
                const int cB = cA + 1;
                F gA[8], gB[8], dA[8], dB[8], pA[8], pB[8];
                if constexpr (RC) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        gA[i] = gR[8 * cA + i];
                        gB[i] = gR[8 * cB + i];
                        dA[i] = dR[8 * cA + i];
                        dB[i] = dR[8 * cB + i];
                        pA[i] = pR[8 * cA + i];
                        pB[i] = pR[8 * cB + i];
                    }
                } else {
                    tmem::ld16(tbase + T_G + CW * cA, gA, gB);
                    tmem::ld16(tbase + T_D + CW * cA, dA, dB);
                    if constexpr (!EURO) {
                        tmem::ld16(tbase + T_P + CW * cA, pA, pB);
                        tmem::hot_wait(gA, gB, dA);
                        tmem::hot_wait(dB, pA, pB);
                    } else {
                        tmem::hot_wait(gA, gB);
                        tmem::hot_wait(dA, dB);
                    }
                }
                auto node = [&](int i, F& uA, F& uB) {
                    uA = fma(gA[i], uA, yA[i]);
                    uB = fma(gB[i], uB, yB[i]);
                    const F rA = fma(dA[i], uA, -vr[8 * cA + i]);
                    const F rB = fma(dB[i], uB, -vr[8 * cB + i]);
                    if constexpr (EURO) {
                        vr[8 * cA + i] = rA;
                        vr[8 * cB + i] = rB;
                    } else {
                        vr[8 * cA + i] = max_like_std(rA, pA[i]);
                        vr[8 * cB + i] = max_like_std(rB, pB[i]);
                    }
                };
                if constexpr (!D4) {
                    F uA = Uin[cA], uB = Uin[cB];
#pragma unroll
                    for (int i = 7; i >= 0; --i) node(i, uA, uB);
                } else {
                    double uhA = Uin[cA], uhB = Uin[cB];
                    double ulA = fma(hQ_hi(cA), Uin[cA], fma(hR_hi(cA), y3[0], ld_l4(cA)));  // u_4* of chunk A
                    double ulB = fma(hQ_hi(cB), Uin[cB], fma(hR_hi(cB), y3[1], ld_l4(cB)));
#pragma unroll
                    for (int i = 3; i >= 0; --i) {
                        node(4 + i, uhA, uhB);
                        node(i, ulA, ulB);
                    }
                }
                F aA[8], aB[8];
                if constexpr (RC) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        aA[i] = aR[8 * cA + i];
                        aB[i] = aR[8 * cB + i];
                    }
                } else {
                    tmem::ld16(tbase2 + T_A + CW * cA, aA, aB);
                    tmem::ld16(tbase2 + T_G + CW * cA, gA, gB);
                    tmem::hot_wait(aA, aB);
                    tmem::hot_wait(gA, gB);
                }
                if constexpr (!D4) {
                    yA[0] = vr[8 * cA];
                    yB[0] = vr[8 * cB];
#pragma unroll
                    for (int i = 1; i < 8; ++i) {
                        yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                        yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                    }
                    e[cA] = yA[7];
                    e[cB] = yB[7];
                    F uA = yA[7], uB = yB[7];
#pragma unroll
                    for (int i = 6; i >= 0; --i) {
                        uA = fma(gA[i], uA, yA[i]);
                        uB = fma(gB[i], uB, yB[i]);
                    }
                    f[cA] = uA;
                    f[cB] = uB;
                } else {
                    // four 3-deep Horner halves forward, four backward, combined through the half-chunk products
                    yA[0] = vr[8 * cA];
                    yB[0] = vr[8 * cB];
                    yA[4] = vr[8 * cA + 4];
                    yB[4] = vr[8 * cB + 4];
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        yA[i] = fma(aA[i], yA[i - 1], vr[8 * cA + i]);
                        yB[i] = fma(aB[i], yB[i - 1], vr[8 * cB + i]);
                        yA[4 + i] = fma(aA[4 + i], yA[3 + i], vr[8 * cA + 4 + i]);
                        yB[4 + i] = fma(aB[4 + i], yB[3 + i], vr[8 * cB + 4 + i]);
                    }
                    e[cA] = fma(hP_hi(cA), yA[3], yA[7]);
                    e[cB] = fma(hP_hi(cB), yB[3], yB[7]);
                    double ulA = yA[3], ulB = yB[3], uhA = yA[7], uhB = yB[7];
#pragma unroll
                    for (int i = 2; i >= 0; --i) {
                        ulA = fma(gA[i], ulA, yA[i]);
                        ulB = fma(gB[i], ulB, yB[i]);
                        uhA = fma(gA[4 + i], uhA, yA[4 + i]);
                        uhB = fma(gB[4 + i], uhB, yB[4 + i]);
                    }
                    f[cA] = fma(hQ_lo(cA), fma(hR_hi(cA), yA[3], uhA), ulA);
                    f[cB] = fma(hQ_lo(cB), fma(hR_hi(cB), yB[3], uhB), ulB);
                    st_e3(cA, yA[3]);
                    st_e3(cB, yB[3]);
                    st_l4(cA, uhA);
                    st_l4(cB, uhB);
                }
            };
            // Rotated loop.  Block X: pair 0 backwards (its forward sweeps were done at the end of the previous
            // iteration).  Block Z: the other pair in full, then the NEXT step's scans with pair 0's forward sweeps
            // behind them -- the shuffles' latency hides behind the sweeps of the same basic block.
            F y0A[8], y0B[8], y30[2] = {F(0), F(0)};
            scan();
            fwd_pair(0, y0A, y0B, y30);
            for (int step = 0; step < nsteps; ++step) {
                if (RC || step < B.opq_lim[0]) back_pair(0, y0A, y0B, y30);  // always true: basic-block boundary
                if (RC || step < B.opq_lim[1]) {
#pragma unroll
                    for (int h = 2; h < NCH; h += 2) {
                        F yA[8], yB[8], y3[2] = {F(0), F(0)};
                        fwd_pair(h, yA, yB, y3);
                        back_pair(h, yA, yB, y3);
                    }
                    scan();
                    fwd_pair(0, y0A, y0B, y30);  // after the last step: computed and dropped
                }
            }
            tmem::wait_ld();  // nothing in flight when the arrays are rewritten
        };
        auto march_levels = [&](auto euro_c) {
            switch (levels) {
                case 1: march(std::integral_constant<int, 1>{}, euro_c); break;
                case 2: march(std::integral_constant<int, 2>{}, euro_c); break;
                case 3: march(std::integral_constant<int, 3>{}, euro_c); break;
                case 4:
                    if constexpr (MAXLEV >= 4) march(std::integral_constant<int, 4>{}, euro_c);
                    break;
                default:
                    if constexpr (MAXLEV >= 5) march(std::integral_constant<int, 5>{}, euro_c);
                    break;
            }
        };
        // ================= epilogue: interpolate every option of this chain ===============================
        auto emit = [&](double* out, bool eu) {
#pragma unroll
            for (int i = 0; i < NODES; ++i) vfin[lane * NODES + i] = vr[i];
            __syncwarp();
            Fd1dBatch Bo = B;
            Bo.prices = out;
            uint32_t q0, q1;
            chain_range(B, my_pde, q0, q1);
            if (!live) q1 = q0;
            const double* vmine = vfin + ph * XT;
            // a long chain is handed to fd1d_long_value_kernel (a CTA per chain) instead of LPP lanes of this warp
            const bool want = B.long_ws && q1 - q0 > KW_LONG_CHAIN && my_pde < 0x80000000u;
            uint32_t slot = 0;
            if (want && pl == 0) slot = atomicAdd(B.long_count, 1u);
            slot = __shfl_sync(FULL, slot, 0, LPP);
            if (want && slot < B.long_cap) {
                double* w = B.long_ws + (size_t)slot * XT;
                for (int j = pl; j < XT; j += LPP) w[j] = vmine[j];
                if (pl == 0) B.long_meta[slot] = my_pde | (eu ? 0x80000000u : 0u);
                q1 = q0;
            }
            for (uint32_t q = q0 + pl; q < q1; q += LPP) {
                const uint32_t oi = B.csr_opt ? __ldg(B.csr_opt + q) : q;
                price_option_sinh_grid(Bo, sc, oi, [&](int j) { return vmine[j]; });
            }
            __syncwarp();  // vfin is rewritten by the next emit
        };
        march_levels(std::false_type{});
        if (pl == 0 && live) {
            // histogram buckets shared with Layout B: 0 = exact requested, 1 = every level, 2/3/4 = 4/3/2, 5 = 1 level
            const int bucket = B.max_mode == 0 ? 0 : (own == MAXLEV ? 1 : 6 - own);
            atomicAdd(&B.status[2 + bucket], 1u);
        }
        emit(B.prices, false);
        if constexpr (BS) {
            if (__any_sync(FULL, sc.american)) {  // (PACK > 1: a PDE given as European marches to the same values again)
                // the European copy: payoff again (src/Pricer/kwFd1d.cpp:127-139), same LU (still in tensor memory), no
                // projection.  An American chain's projection floor IS its payoff on the nodes 0 .. xDim-2 and still sits in
                // tensor memory: only the last node's payoff is evaluated again (and, in a packed warp, the payoff of a PDE given
                // as European, whose floor is "never": it marches to the same values again).
                const bool no_floor = !sc.american;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    F fl[8];
                    tmem::ld8(tbase + T_P + CW * c, fl);
                    tmem::wait_ld_dep(fl);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int j = pl * NODES + 8 * c + i;
                        double p = (double)fl[i];
                        if (j >= xDim) p = 0.;
                        else if (no_floor || j == xDim - 1) p = payoff_node_ni(sc.put, x_node_ni(sc, B.density, j));
                        vfin[(8 * c + i) * 32 + lane] = p;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < NODES; ++i) vr[i] = (F)vfin[i * 32 + lane];
                __syncwarp();
                march_levels(std::true_type{});
            }
            emit(B.prices_eu, true);  // a chain given as European: one march, both arrays
        }
        __syncwarp();  // vfin is rewritten by the next PDE
    }
    tmem::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem::dealloc<64 * NCH>(s_taddr);
}

}  // namespace kwfd1d
