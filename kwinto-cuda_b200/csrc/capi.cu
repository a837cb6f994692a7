// capi.cu -- the extern "C" layer of include/kw_fd1d.h: handle, device-resident asset/price
// buffers, host-side chain compression, kernel dispatch.
//
// Reference functions replaced on this path:
//   Fd1d_Pricer::init   src/Pricer/kwFd1d.cpp:9-19     -> kw_fd1d_create
//   Fd1d_Pricer::price  src/Pricer/kwFd1d.cpp:21-160   -> kw_fd1d_price / kw_fd1d_price_device
//   Fd1d_BlackScholes_Pricer::price src/Pricer/kwFd1d_BlackScholes.cpp:15-43 -> kw_fd1d_price_bs
// There is NO CPU fallback here: without a CUDA device every entry point fails loudly.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "fd1d_common.cuh"
#include "fd1d_reg.cuh"
#include "fd1d_soa.cuh"
#include "fd1d_warp.cuh"
#include "fd1d_iw.cuh"
#include "fd1d_wide.cuh"
#include "fd1d_warpf.cuh"
#ifdef KW_EXPERIMENTS
#include "fd1d_warp_bs.cuh"
#endif
#include "compress.cuh"
#include "microbench.cuh"

using namespace kwfd1d;

namespace {

// ---------------------------------------------------------------- variants of Layout B
typedef void (*RegKernel)(const Fd1dBatch);
typedef void (*WideSetup)(const Fd1dBatch, double*, uint32_t, uint32_t, int);
typedef void (*WideMarch)(const Fd1dBatch, const double*, uint32_t, uint32_t);
struct RegVariant {
    int id;        // 100*log2(P/32) + serial (+1000 for the fp32 march); see the table
    int prec;      // KW_FD1D_F64 / KW_FD1D_F32
    int M, P, minb;
    bool proj_smem, dq_smem;
    RegKernel fn;
    size_t smem;
    int tmem_cols;  // tensor-memory columns each CTA allocates (0 = none)
    int pdes_per_cta;  // 0/1: one PDE per CTA (Layout B); 4: one PDE per warp (Layout W)
    // wide Layout W (fd1d_wide.cuh): NWP warps per PDE, set-up in a kernel of its own through an HBM workspace
    int wide_nwp;
    WideSetup setup_fn;
    WideMarch wide_fn;
    size_t setup_smem;
    size_t slot_doubles;
    int icmp;
};

#define KW_VARIANT(ID, M_, P_, MINB_, PJ_, DQ_)                                                   \
    {                                                                                              \
        ID, KW_FD1D_F64, M_, P_, MINB_, PJ_, DQ_, fd1d_reg_kernel<double, M_, P_, MINB_, PJ_, DQ_>,  \
            RegSmem<M_, P_>::bytes(PJ_, DQ_)                                                       \
    }
#define KW_VARIANT_I(ID, M_, P_, MINB_, PJ_, DQ_)                                                 \
    {                                                                                              \
        ID, KW_FD1D_F64, M_, P_, MINB_, PJ_, DQ_, fd1d_reg_kernel<double, M_, P_, MINB_, PJ_, DQ_, true>, \
            RegSmem<M_, P_>::bytes(PJ_, DQ_)                                                       \
    }

#define KW_VARIANT_T(ID, MINB_, ICMP_)                                                            \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 128, MINB_, false, false, fd1d_reg_kernel<double, 8, 128, MINB_, false, false, ICMP_, true>, \
            RegSmem<8, 128>::bytes(false, false), 128                                              \
    }
#define KW_VARIANT_W(ID, MINB_, ICMP_, PAIR_)                                                     \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 128, MINB_, false, false, fd1d_warp_kernel<4, MINB_, ICMP_, PAIR_>,    \
            WarpSmem<4>::bytes(), 256, 4                                                           \
    }
#define KW_VARIANT_WS(ID, NCH_, MINB_)                                                            \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 32 * NCH_, MINB_, false, false, fd1d_warp_kernel<NCH_, MINB_, false, true, 0, false, true>, \
            WarpSmem<NCH_>::bytes(), 64 * NCH_, 4                                                  \
    }
#define KW_VARIANT_IW(ID, NCH_, MINB_)                                                            \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 32 * NCH_, MINB_, false, false, fd1d_iw_kernel<NCH_, MINB_>,           \
            IwSmem<NCH_>::bytes(), 64 * NCH_, 4                                                    \
    }
#define KW_VARIANT_IWP(ID, PACK_)                                                                 \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 128 / PACK_, 2, false, false, fd1d_iw_kernel<4, 2, false, false, PACK_>, \
            IwSmem<4>::bytes(), 256, 4 * PACK_                                                     \
    }
#define KW_VARIANT_IWF(ID, PACK_)                                                                 \
    {                                                                                              \
        ID, KW_FD1D_F32, 8, 128 / PACK_, 2, false, false, fd1d_iw_kernel<4, 2, false, false, PACK_, float>, \
            IwSmem<4>::bytes(), 256, 4 * PACK_                                                     \
    }
#define KW_VARIANT_WRT(ID, MINB_)                                                                 \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 128, MINB_, false, false, fd1d_warp_kernel<4, MINB_, false, true, 0, true>, \
            WarpSmem<4>::bytes(), 256, 4                                                           \
    }
#define KW_VARIANT_WN(ID, NCH_, MINB_)                                                            \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 32 * NCH_, MINB_, false, false, fd1d_warp_kernel<NCH_, MINB_, false, true>, \
            WarpSmem<NCH_>::bytes(), 64 * NCH_, 4                                                  \
    }
#define KW_VARIANT_WF(ID, NCH_, MINB_)                                                            \
    {                                                                                              \
        ID, KW_FD1D_F32, 8, 32 * NCH_, MINB_, false, false, fd1d_warpf_kernel<NCH_, MINB_, false, true, float>, \
            WarpSmem<NCH_>::bytes(), 32 * NCH_, 4                                                  \
    }
#define KW_VARIANT_W2(ID, MINB_, ICMP_)                                                           \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 128, MINB_, false, false, fd1d_warp2_kernel<4, MINB_, ICMP_>,          \
            Warp2Smem<4>::bytes(), 256, 4                                                          \
    }
#define KW_VARIANT_WIDE(ID, NWP_, ICMP_)                                                          \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 128 * NWP_, 2, false, false, nullptr, WideSmem<NWP_>::bytes(), 256,    \
            4 / NWP_, NWP_, fd1d_wide_setup_kernel<128 * NWP_>, fd1d_wide_kernel<NWP_, 2, ICMP_>,  \
            sizeof(double) * 16 * 128 * NWP_, WideSlot<128 * NWP_>::doubles, ICMP_                 \
    }
#define KW_VARIANT_WIDES(ID, NWP_)                                                                \
    {                                                                                              \
        ID, KW_FD1D_F64, 8, 128 * NWP_, 2, false, false, nullptr, WideSmem<NWP_>::bytes(), 256,    \
            4 / NWP_, NWP_, fd1d_wide_setup_kernel<128 * NWP_>, fd1d_wide_kernel<NWP_, 2, false, true>, \
            sizeof(double) * 16 * 128 * NWP_, WideSlot<128 * NWP_>::doubles, false                 \
    }
#define KW_VARIANT_F32(ID, M_, P_, MINB_)                                                         \
    {                                                                                              \
        ID, KW_FD1D_F32, M_, P_, MINB_, false, false, fd1d_reg_kernel<float, M_, P_, MINB_, false, false>, \
            RegSmem<M_, P_>::bytes(false, false)                                                   \
    }

// The first entry of each P is the default of the per-xDim dispatch.  The shipped library carries the
// kernels the dispatch can reach; every other variant measured along the way (DESIGN.md "Tried") is compiled only
// with -DKW_EXPERIMENTS (make EXPERIMENTS=1), so that they cost no build time, binary size or test matrix by default.
const RegVariant g_variants[] = {
    KW_VARIANT_IWP(38, 4),                     // x <= 256: 237 with four PDEs per warp (8 lanes each; 1.87 vs 2.31 ms at 256^2)
    KW_VARIANT(1, 8, 32, 12, false, false),    // x <= 256, CTA per PDE (small batches)
    KW_VARIANT_IWP(138, 2),                    // x <= 512: 237 with two PDEs per warp (16 lanes each; 6.36 vs 7.03 ms at 512^2)
    KW_VARIANT(101, 8, 64, 6, false, false),   // x <= 512, CTA per PDE (small batches)
    KW_VARIANT_IW(237, 4, 2),                  // x <= 1024: Layout W with independent warps (fd1d_iw.cuh): warp-level set-up, no CTA
                                               // barrier, PDEs handed out by an atomic counter, rotated split march (23.9 ms)
    KW_VARIANT(201, 8, 128, 3, false, false),  // x <= 1024, CTA per PDE (batches below one wave of Layout W)
    KW_VARIANT_WIDES(336, 2),                  // x <= 2048: Layout W over two warps per PDE, rotated split march (15.65 vs 17.7 ms)
    KW_VARIANT(301, 8, 256, 1, false, false),  // x <= 2048, CTA per PDE (small batches)
    KW_VARIANT_WIDES(436, 4),                  // x <= 4096: Layout W over four warps per PDE, rotated split march (31.3 vs 36.0 ms)
    KW_VARIANT(401, 8, 512, 1, true, true),    // x <= 4096, CTA per PDE (small batches)
    // fp32 march (fp64 set-up): FD1D.GPU.PRECISION = f32
    // Layout W: the independent-warp kernel with a float march, every coefficient of the lane in registers (fd1d_iw.cuh, RC):
    // 1024^2 12.55 ms against 14.9 ms for round 1's 1233, 512^2 3.73 against 4.51 ms (1101), 256^2 1.01 against 1.35 ms (1001)
    KW_VARIANT_IWF(1038, 4),     // x <= 256, four PDEs per warp
    KW_VARIANT_F32(1001, 8, 32, 16),
    KW_VARIANT_IWF(1138, 2),     // x <= 512, two PDEs per warp
    KW_VARIANT_F32(1101, 8, 64, 8),
    KW_VARIANT_IWF(1237, 1),     // x <= 1024
    KW_VARIANT_F32(1201, 8, 128, 4),
    KW_VARIANT_F32(1301, 8, 256, 2),
    KW_VARIANT_F32(1401, 8, 512, 1),
#ifdef KW_EXPERIMENTS
    {239, KW_FD1D_F64, 8, 128, 2, false, false, fd1d_iw_kernel<4, 2, false, true>, IwSmem<4>::bytes(), 256, 4},  // 237 with half-chunk
                                               // lookahead (four dependent chains per warp, +16 DFMAs per step): 28.4 vs 24.1 ms
    KW_VARIANT(2, 8, 32, 16, true, true),
    KW_VARIANT(102, 8, 64, 8, true, true),
    KW_VARIANT(103, 8, 64, 6, true, false),
    KW_VARIANT_WS(236, 4, 2),                  // Layout W, CTA-cooperative set-up, chunk pairs, every pair phase a basic block of its
                                               // own, a~ and g~ loaded twice (25.8 vs 27.6 ms for 233)
    KW_VARIANT_W(233, 2, false, true),         // the round-1 default: the chunk-pair phases of a step in one basic block
    KW_VARIANT(202, 8, 128, 3, true, false),
    KW_VARIANT(203, 8, 128, 4, true, true),
    KW_VARIANT(204, 8, 128, 4, true, false),
    KW_VARIANT(205, 8, 128, 3, true, true),
    KW_VARIANT_I(211, 8, 128, 3, false, false),  // 201 with the projection compare on the integer pipe
    KW_VARIANT_I(213, 8, 128, 4, true, true),
    KW_VARIANT_T(221, 4, false),  // coefficient arrays in tensor memory, 4 PDEs per SM
    KW_VARIANT_T(222, 4, true),
    KW_VARIANT_W(232, 2, true, false),
    KW_VARIANT_W(231, 2, false, false),  // one chunk at a time, next chunk's a~ prefetched
    KW_VARIANT_W(234, 2, true, true),
    KW_VARIANT_WN(133, 2, 2),  // the round-1 default for x <= 512: Layout W with 2 chunks per lane (one chunk pair)
    KW_VARIANT_IW(137, 2, 2),  // 237's two-chunk twin (7.30 ms at 512^2 against 6.91 for 133)
    KW_VARIANT_WS(136, 2, 2),  // 133 in the form of 236 (slower on the two-chunk tile: 7.65 vs 6.92 ms at 512^2)
    KW_VARIANT_WRT(235, 2),  // 233 with the scan-level count as a run-time value: one march loop instead of five
    KW_VARIANT_W2(241, 2, false),  // v in tensor memory, floor from shared memory
    KW_VARIANT_W2(242, 2, true),
    KW_VARIANT_WIDE(331, 2, false),            // the round-1 two-warp kernel
    KW_VARIANT(302, 8, 256, 2, true, true),
    KW_VARIANT_WIDE(431, 4, false),            // the round-1 four-warp kernel
    KW_VARIANT(402, 8, 512, 1, true, false),
    KW_VARIANT_WF(1233, 4, 2),   // round 1's Layout W with a float march, 512 < x <= 1024 (CTA-cooperative set-up, coefficients in tensor memory)
    KW_VARIANT_WF(1133, 2, 2),   // Layout W, float march, 256 < x <= 512 (slower than 1101: 6.0 M vs 7.1 M options/s)
#endif
};
// fused FD1D-BS marches (one set-up and one tensor-memory copy of a~, g~, D for the solve as given and the
// solve of the European copy).  257 / 158 / 58 (fd1d_iw.cuh, BS; one, two, four PDEs per warp), default where they apply: every
// warp marches its chain as given, then the European copy.  Experiments: 252 (FD1D.GPU.BS_FUSED = 3; BS = 1): eight warps per CTA,
// warp w marches PDE w as given while warp w + 4 marches the European copy; 251 (FD1D.GPU.BS_FUSED = 2;
// fd1d_warp_bs.cuh): both solutions in one warp's step (instruction-cache bound, slower than two solves).
const RegVariant g_bs2_variant = {257, KW_FD1D_F64, 8, 128, 2, false, false, fd1d_iw_kernel<4, 2, true>,
                                  IwSmem<4>::bytes(), 256, 4};  // 512 < x <= 1024: the independent-warp kernel, BS = true
const RegVariant g_bs2n2_variant = {158, KW_FD1D_F64, 8, 64, 2, false, false, fd1d_iw_kernel<4, 2, true, false, 2>,
                                    IwSmem<4>::bytes(), 256, 8};  // 256 < x <= 512: two PDEs per warp
const RegVariant g_bs2n1_variant = {58, KW_FD1D_F64, 8, 32, 2, false, false, fd1d_iw_kernel<4, 2, true, false, 4>,
                                    IwSmem<4>::bytes(), 256, 16};  // x <= 256: four PDEs per warp
// the same three with the fp32 march (FD1D.GPU.PRECISION = f32)
const RegVariant g_bsf_variant[3] = {
    {1257, KW_FD1D_F32, 8, 128, 2, false, false, fd1d_iw_kernel<4, 2, true, false, 1, float>, IwSmem<4>::bytes(), 256, 4},
    {1158, KW_FD1D_F32, 8, 64, 2, false, false, fd1d_iw_kernel<4, 2, true, false, 2, float>, IwSmem<4>::bytes(), 256, 8},
    {1058, KW_FD1D_F32, 8, 32, 2, false, false, fd1d_iw_kernel<4, 2, true, false, 4, float>, IwSmem<4>::bytes(), 256, 16}};
// 1024 < x <= 2048 / 4096: the wide kernels with BS = true
const RegVariant g_bs_wide2_variant = {356, KW_FD1D_F64, 8, 256, 2, false, false, nullptr, WideSmem<2>::bytes(), 256, 2, 2,
                                       fd1d_wide_setup_kernel<256>, fd1d_wide_kernel<2, 2, false, true, true>,
                                       sizeof(double) * 16 * 256, WideSlot<256>::doubles, false};
const RegVariant g_bs_wide4_variant = {456, KW_FD1D_F64, 8, 512, 2, false, false, nullptr, WideSmem<4>::bytes(), 256, 1, 4,
                                       fd1d_wide_setup_kernel<512>, fd1d_wide_kernel<4, 2, false, true, true>,
                                       sizeof(double) * 16 * 512, WideSlot<512>::doubles, false};
#ifdef KW_EXPERIMENTS
const RegVariant g_bs253_variant = {253, KW_FD1D_F64, 8, 128, 2, false, false, fd1d_warp_kernel<4, 2, false, true, 2, true>,
                                    WarpSmem<4>::bytes(), 256, 4};  // round 1's fused kernel (CTA-cooperative set-up)
const RegVariant g_bs_variant = {252, KW_FD1D_F64, 8, 256, 1, false, false, fd1d_warp_kernel<4, 1, false, true, 1, true>,
                                 WarpSmem<4, 256>::bytes(), 256, 4};
const RegVariant g_bs1_variant = {251, KW_FD1D_F64, 8, 128, 2, false, false, fd1d_warp_bs_kernel<2>,
                                  Warp2Smem<4>::bytes(), 256, 4};
#endif
constexpr int kNumVariants = sizeof(g_variants) / sizeof(g_variants[0]);
constexpr int kMaxRegX = 4096;

const RegVariant* find_variant(int xDim, int want_id, int prec)
{
    const RegVariant* first_fit = nullptr;
    for (int i = 0; i < kNumVariants; ++i) {
        const RegVariant& v = g_variants[i];
        if (v.M * v.P < xDim || v.prec != prec) continue;
        if (!first_fit || v.P < first_fit->P) first_fit = &v;
    }
    if (!first_fit) return nullptr;
    if (want_id > 0) {
        for (int i = 0; i < kNumVariants; ++i)
            if (g_variants[i].id == want_id && g_variants[i].prec == prec && g_variants[i].M * g_variants[i].P >= xDim)
                return &g_variants[i];
        return nullptr;
    }
    // auto: the first entry of the smallest fitting P (the table lists the default first)
    for (int i = 0; i < kNumVariants; ++i)
        if (g_variants[i].P == first_fit->P && g_variants[i].prec == prec) return &g_variants[i];
    return first_fit;
}

// Before every march launch.  `keep_errors`: an earlier batch of this handle has not been synchronised yet
// (kw_fd1d_price_device called several times before one kw_fd1d_sync): its range-error count and first failing
// index must survive until the sync reads them, so only the per-launch fields are cleared.
__global__ void status_reset_kernel(unsigned int* status, int keep_errors)
{
    if (!keep_errors) {
        status[0] = 0u;
        status[1] = 0xffffffffu;
        status[2] = status[3] = status[4] = status[5] = status[6] = status[7] = 0u;
    }
    status[8] = 0u;  // work counter of the independent-warp kernels
    status[9] = 0u;  // long-chain slots handed out
}

// device-side chain compression of `n` device-resident options; fills the batch's PDE tables
int compress_on_device(kw_fd1d_handle* h, Fd1dBatch& B, const kw_option* d_opts, size_t n, cudaStream_t st);

// BlackScholes_Pricer::priceOne (src/Pricer/kwBlackScholes.cpp:27-50) on European copies and
// the control-variate combination of src/Pricer/kwFd1d_BlackScholes.cpp:38-40:
//   prices[i] += bs[i] - fdEuro[i]
__global__ void bs_combine_kernel(const kw_option* opts, size_t n, double* prices, const double* fd_euro)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const kw_option o = load_option(opts + i);
    const double w = (double)o.w;
    const double zt = __dmul_rn(o.z, sqrt(o.t));
    const double drift = __dadd_rn(__dsub_rn(o.r, o.q), __dmul_rn(__dmul_rn(0.5, o.z), o.z));
    const double d1 = __dmul_rn(1 / zt, __dadd_rn(log(o.s / o.k), __dmul_rn(drift, o.t)));
    const double d2 = __dsub_rn(d1, zt);
    const double isq2 = 1.4142135623730951;
    const double n1 = 0.5 * (1 + erf(w * d1 / isq2));
    const double n2 = 0.5 * (1 + erf(w * d2 / isq2));
    const double bs = w * (__dsub_rn(__dmul_rn(__dmul_rn(o.s, n1), exp(-o.q * o.t)),
                                     __dmul_rn(__dmul_rn(o.k, n2), exp(-o.r * o.t))));
    prices[i] = __dadd_rn(prices[i], __dsub_rn(bs, fd_euro[i]));
}

struct KeyHash {
    size_t operator()(const kw_option& o) const
    {
        auto mix = [](uint64_t h, uint64_t v) {
            h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
            return h;
        };
        uint64_t b[4];
        const double f[4] = {o.t + 0.0, o.r + 0.0, o.q + 0.0, o.z + 0.0};  // -0.0 == +0.0 in KeyEq: hash them alike
        memcpy(b, f, sizeof b);
        uint64_t h = 0x243f6a8885a308d3ull;
        for (int i = 0; i < 4; ++i) h = mix(h, b[i] * 0xff51afd7ed558ccdull);
        h = mix(h, ((uint64_t)o.e << 8) | (uint8_t)o.w);
        return (size_t)h;
    }
};
struct KeyEq {
    // the reference's chain key (src/Pricer/kwFd1d.cpp:33-35): (t, r, q, z, e, w); s, k are not in it
    bool operator()(const kw_option& l, const kw_option& r) const
    {
        return l.t == r.t && l.r == r.r && l.q == r.q && l.z == r.z && l.e == r.e && l.w == r.w;
    }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct kw_fd1d_handle {
    kw_fd1d_config cfg;
    std::string err;
    int sm_count = 0;
    int clock_khz = 0;
    char name[128] = {0};

    int layout = 0;
    const RegVariant* var = nullptr;        // variant for batches that fill the device
    int ctas_per_sm = 0;
    int regs = 0;
    const RegVariant* var_small = nullptr;  // auto dispatch only: variant for batches below `small_below` PDEs
    int ctas_per_sm_small = 0;
    int regs_small = 0;
    uint32_t small_below = 0;
    const RegVariant* var_bs = nullptr;     // fused FD1D-BS march, when the configuration has one
    int ctas_per_sm_bs = 0;
    int regs_bs = 0;
    bool bs_forced = false;                 // FD1D.GPU.BS_FUSED = 2 / 3: fused for every batch size
    const RegVariant* last_var = nullptr;   // what the last batch ran
    uint32_t class_split = 0;               // != 0: the last batch launched var_small AND var; chains < class_split ran var_small
    int last_grid = 0;
    int launches = 0;  // kernels launched by the current / last price call
    uint64_t last_n_pde = 0;
    unsigned int mode_count[6] = {0, 0, 0, 0, 0, 0};
    unsigned int long_chains = 0;           // slots fd1d_long_value_kernel priced in the last synchronised call

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;

    DevBuf<kw_option> d_opts, d_opts2;
    DevBuf<double> d_prices, d_prices2;
    DevBuf<uint32_t> d_rep, d_start, d_csr;
    DevBuf<double> d_ws;       // wide Layout W: set-up workspace for one chunk of the batch
    DevBuf<uint32_t> d_chain;  // device-side compression: one slab carved into the ChainTable arrays
    bool dev_compressed = false;
    DevBuf<unsigned int> d_status;
    DevBuf<double> d_long;        // final v of the chains with more than KW_LONG_CHAIN options (fd1d_long_value_kernel)
    DevBuf<uint32_t> d_long_meta;
    DevBuf<double> d_soa;
    PinBuf<uint32_t> h_idx;  // rep | start | csr staging
    PinBuf<unsigned int> h_status;

    // Multi-device handle (kw_fd1d_create_multi / cfg.n_devices > 1): this object only coordinates; every entry
    // of `shards` is a complete single-device handle (its own stream, device buffers, pinned staging) and prices a
    // contiguous block of the batch on its device.  Empty for a single-device handle.
    bool unsynced = false;  // a batch was enqueued and its status not read yet (check_status clears)
    std::vector<kw_fd1d_handle*> shards;
    uint32_t shards_used = 0;  // shards the last call spread the batch over
    double last_wall_ms = 0.;  // multi-device: wall time of the last price call (max over shards is in last_kernel_ms)
};

namespace {

int fail(kw_fd1d_handle* h, int code, const std::string& msg)
{
    if (h) h->err = msg;
    return code;
}

#define KW_CUDA(h, call)                                                                       \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(h, KW_FD1D_ECUDA,                                                      \
                        std::string("Fd1dGpu_Pricer: CUDA error: ") + cudaGetErrorString(e_) + \
                            " at " #call);                                                     \
    } while (0)

int prepare_variant(kw_fd1d_handle* h, const RegVariant* var, const cudaDeviceProp& prop, int& ctas_per_sm, int& regs);
const RegVariant* find_small_variant(int xDim, int prec);

// Layout A streams 7 arrays of xDim doubles per PDE through HBM and has ONE thread per PDE: it needs hundreds of
// thousands of PDEs in flight to hide the latency of its dependent sweeps, so a chunk is as many PDEs as fit half of
// the free device memory (180 GB hold 1.6 M PDEs at x = 1024), not a fixed 65536.
size_t soa_chunk(const kw_fd1d_handle* h, size_t n_pde)
{
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)8 << 30;
    const size_t per_pde = 7 * sizeof(double) * (size_t)h->cfg.x_grid_size;
    size_t fit = std::max<size_t>(4096, (free_b / 2 + h->d_soa.cap * sizeof(double)) / per_pde);
    if (const char* e = getenv("KW_FD1D_SOA_CHUNK")) fit = std::max<size_t>(128, (size_t)atoll(e));
    return std::min<size_t>(n_pde, fit);
}

// enqueue the solve of `n_pde` PDEs on `st`; all pointers are device pointers
int launch_batch(kw_fd1d_handle* h, Fd1dBatch B, cudaStream_t st)
{
    for (int i = 0; i < 4; ++i) B.opq_lim[i] = INT32_MAX;
    B.opq_zero = 0;
    B.pde_lo = 0;
    B.pde_hi = 0xffffffffu;
    B.work_counter = B.status + 8;
    status_reset_kernel<<<1, 1, 0, st>>>(B.status, h->unsynced ? 1 : 0);
    h->unsynced = true;
    h->launches += 1;
    h->last_n_pde = B.n_pde;
    if (h->layout == KW_FD1D_LAYOUT_REG) {
        const bool fused = B.prices_eu != nullptr;
        if (fused && !h->var_bs) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: no fused FD1D-BS kernel for this configuration");
        // one variant over the batch `Bx` (its PDE-count class [pde_lo, pde_hi) decides on the device whether it runs at all)
        auto run = [&](const RegVariant* v, int ctas_per_sm, const Fd1dBatch& Bx, uint32_t n_max) -> cudaError_t {
            const uint32_t ppc = v->pdes_per_cta > 1 ? (uint32_t)v->pdes_per_cta : 1u;
            if (v->wide_nwp) {
                // set-up kernel -> HBM workspace -> march kernel, one chunk of the batch at a time
                // chunk = as many PDEs as fit a 2 GiB workspace (a multiple of one wave of the march grid)
                const uint32_t wave = (uint32_t)(h->sm_count * ctas_per_sm) * ppc;
                uint32_t fit = (uint32_t)(((size_t)2 << 30) / (v->slot_doubles * sizeof(double)));
                fit = std::max(wave, fit / wave * wave);
                const uint32_t cap = std::min<uint32_t>(n_max, fit);
                if (cudaError_t e = h->d_ws.reserve((size_t)cap * v->slot_doubles)) return e;
                const int P = 128 * v->wide_nwp;
                for (uint32_t base = 0; base < n_max; base += cap) {
                    const uint32_t cnt = std::min<uint32_t>(cap, n_max - base);
                    const int gs = (int)std::min<uint32_t>(cnt, (uint32_t)(h->sm_count * (2048 / P)));
                    v->setup_fn<<<gs, P, v->setup_smem, st>>>(Bx, h->d_ws.p, base, cnt, v->icmp);
                    const uint32_t wantc = (cnt + ppc - 1) / ppc;
                    const int gm = (int)std::min<uint32_t>(wantc, (uint32_t)(h->sm_count * ctas_per_sm));
                    v->wide_fn<<<gm, 128, v->smem, st>>>(Bx, h->d_ws.p, base, cnt);
                    h->launches += 2;
                    h->last_grid = gm;
                }
                return cudaSuccess;
            }
            int grid = h->sm_count * ctas_per_sm;
            const uint32_t want = (n_max + ppc - 1) / ppc;
            if ((uint32_t)grid > want) grid = (int)want;
            h->last_grid = grid;
            v->fn<<<grid, v->pdes_per_cta > 1 ? std::max(128, v->P) : v->P, v->smem, st>>>(Bx);
            h->launches += 1;
            return cudaSuccess;
        };
        // long chains of the independent-warp kernels: slots for the final v, priced by a kernel of their own afterwards
        const RegVariant* vbig = fused ? h->var_bs : h->var;
        const bool iw = vbig->fn && vbig->pdes_per_cta >= 4 && vbig->tmem_cols == 256 && B.csr_start && B.n_opt > KW_LONG_CHAIN &&
                        !getenv("KW_FD1D_NO_LONG");  // (the variable exists for tools/long_chain_probe.py's comparison)
        const int tile = vbig->M * vbig->P;
        if (iw) {
            B.long_cap = (B.n_opt / KW_LONG_CHAIN + 1) * (fused ? 2u : 1u);
            KW_CUDA(h, h->d_long.reserve((size_t)B.long_cap * tile));
            KW_CUDA(h, h->d_long_meta.reserve(B.long_cap));
            B.long_ws = h->d_long.p;
            B.long_meta = h->d_long_meta.p;
            B.long_count = B.status + 9;
        }
        KW_CUDA(h, cudaEventRecord(h->ev0, st));
        h->class_split = 0;
        if (!fused && h->var_small && B.n_pde_dev && B.n_pde >= h->small_below) {
            // Device-side chain compression: B.n_pde is the OPTION count, the number of chains is known only on the device, and the
            // dispatch is by chains (a portfolio of 6000 options in 600 chains belongs to the small-batch kernel).  Reading the
            // count back would cost a stream synchronisation in the middle of an asynchronous call; instead BOTH kernels are
            // launched and each returns at once unless the count falls into its class (batch_n_pde(), fd1d_common.cuh): an
            // idle launch costs 2-3 us.  kw_fd1d_sync reports which one ran.
            Fd1dBatch Bs = B, Bb = B;
            Bs.pde_hi = h->small_below;
            Bb.pde_lo = h->small_below;
            KW_CUDA(h, run(h->var_small, h->ctas_per_sm_small, Bs, h->small_below - 1));
            KW_CUDA(h, run(h->var, h->ctas_per_sm, Bb, B.n_pde));
            h->class_split = h->small_below;
            h->last_var = h->var;
        } else {
            const bool small = !fused && h->var_small && B.n_pde < h->small_below;
            const RegVariant* v = fused ? h->var_bs : (small ? h->var_small : h->var);
            h->last_var = v;
            KW_CUDA(h, run(v, fused ? h->ctas_per_sm_bs : (small ? h->ctas_per_sm_small : h->ctas_per_sm), B, B.n_pde));
        }
        if (iw) {
            fd1d_long_value_kernel<<<(int)std::min<uint32_t>(B.long_cap, (uint32_t)(4 * h->sm_count)), 256, 0, st>>>(B, tile);
            h->launches += 1;
        }
        KW_CUDA(h, cudaEventRecord(h->ev1, st));
        h->ev_valid = true;
        KW_CUDA(h, cudaGetLastError());
        return KW_FD1D_OK;
    }
    // Layout A: chunks of PDEs, 7 SoA arrays each
    const size_t chunk = soa_chunk(h, B.n_pde);
    const size_t per = (size_t)B.xDim * chunk;
    KW_CUDA(h, h->d_soa.reserve(7 * per));
    KW_CUDA(h, cudaEventRecord(h->ev0, st));
    for (size_t base = 0; base < B.n_pde; base += chunk) {
        const uint32_t cnt = (uint32_t)std::min<size_t>(chunk, B.n_pde - base);
        SoaWork W;
        const size_t pitch = (size_t)B.xDim * cnt;
        W.A = h->d_soa.p;
        W.G = W.A + pitch;
        W.D = W.G + pitch;
        W.PR = W.D + pitch;
        W.V = W.PR + pitch;
        W.Y = W.V + pitch;
        W.X = W.Y + pitch;
        W.n = cnt;
        Fd1dBatch Bc = B;
        Bc.pde_base = (uint32_t)base;
        const int tpb = 128;
        const int grid = (int)((cnt + tpb - 1) / tpb);
        h->last_grid = grid;
        fd1d_soa_setup_kernel<<<grid, tpb, 0, st>>>(Bc, W);
        fd1d_soa_march_kernel<<<grid, tpb, 0, st>>>(Bc, W);
        fd1d_soa_value_kernel<<<grid, tpb, 0, st>>>(Bc, W);
        h->launches += 3;
    }
    KW_CUDA(h, cudaEventRecord(h->ev1, st));
    h->ev_valid = true;
    KW_CUDA(h, cudaGetLastError());
    return KW_FD1D_OK;
}

// host mirror of the reference's error text (src/Math/kwFd1d.cpp:151-153 via
// src/Pricer/kwFd1d.cpp:154-155)
std::string range_message(const kw_fd1d_handle* h, const kw_option& o)
{
    const double x_ = std::log(o.s / o.k);
    const double half = h->cfg.scale * o.z * std::sqrt(o.t);
    const double x0 = h->cfg.density * std::sinh(std::asinh((0. - half) / h->cfg.density));
    const double x1 = h->cfg.density * std::sinh(std::asinh((0. + half) / h->cfg.density));
    return "Fd1d_Pricer::price Fd1d::value: x=" + std::to_string(x_) + " not in range (" +
           std::to_string(x0) + ", " + std::to_string(x1) + ")";
}

// Chain compression (src/Pricer/kwFd1d.cpp:28-65): rep[m], start[m+1], csr[n] into pinned staging.
// A hash on the key replaces the reference's sort; PDE numbering differs, prices do not.
int compress(kw_fd1d_handle* h, const kw_option* a, size_t n, size_t& m, uint32_t*& rep, uint32_t*& start,
             uint32_t*& csr)
{
    KW_CUDA(h, h->h_idx.reserve(3 * n + 2));
    rep = h->h_idx.p;
    start = rep + n;
    csr = start + n + 1;
    std::vector<uint32_t> a2p(n);
    std::unordered_map<kw_option, uint32_t, KeyHash, KeyEq> map;
    map.reserve(n * 2);
    m = 0;
    for (size_t i = 0; i < n; ++i) {
        auto it = map.find(a[i]);
        if (it == map.end()) {
            map.emplace(a[i], (uint32_t)m);
            rep[m] = (uint32_t)i;
            a2p[i] = (uint32_t)m;
            ++m;
        } else {
            a2p[i] = it->second;
        }
    }
    std::fill(start, start + m + 1, 0u);
    for (size_t i = 0; i < n; ++i) start[a2p[i] + 1]++;
    for (size_t p = 0; p < m; ++p) start[p + 1] += start[p];
    std::vector<uint32_t> fillp(start, start + m);
    for (size_t i = 0; i < n; ++i) csr[fillp[a2p[i]]++] = (uint32_t)i;
    return KW_FD1D_OK;
}

int compress_on_device(kw_fd1d_handle* h, Fd1dBatch& B, const kw_option* d_opts, size_t n, cudaStream_t st)
{
    uint64_t cap64 = 64;
    while (cap64 < 2 * (uint64_t)n) cap64 <<= 1;
    if (cap64 > (1ull << 31))
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: batch too large for device-side chain compression (FD1D.GPU.COMPRESS = 1 takes up to 2^30 options)");
    const uint32_t cap = (uint32_t)cap64;
    // slab: slot_rep, slot_cnt, slot_pde [cap] | opt_slot, rep, seg_start, seg_cnt, seg_fill, members [n] | counters [2]
    KW_CUDA(h, h->d_chain.reserve(3 * (size_t)cap + 6 * n + 2));
    ChainTable T;
    uint32_t* p = h->d_chain.p;
    T.slot_rep = p;
    T.slot_cnt = p + cap;
    T.slot_pde = p + 2 * (size_t)cap;
    p += 3 * (size_t)cap;
    T.opt_slot = p;
    T.rep = p + n;
    T.seg_start = p + 2 * n;
    T.seg_cnt = p + 3 * n;
    T.seg_fill = p + 4 * n;
    T.members = p + 5 * n;
    T.counters = p + 6 * n;
    T.cap = cap;
    const unsigned tpb = 256;
    chain_reset_kernel<<<(cap + tpb - 1) / tpb, tpb, 0, st>>>(T);
    chain_insert_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, st>>>(d_opts, (uint32_t)n, T);
    chain_compact_kernel<<<(cap + tpb - 1) / tpb, tpb, 0, st>>>(T);
    chain_fill_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, st>>>((uint32_t)n, T);
    h->launches += 4;
    KW_CUDA(h, cudaGetLastError());
    B.pde_rep = T.rep;
    B.csr_start = T.seg_start;
    B.csr_cnt = T.seg_cnt;
    B.csr_opt = T.members;
    B.n_pde_dev = T.counters;
    B.n_pde = (uint32_t)n;  // upper bound (sizes the grid); the kernels read the true count
    h->dev_compressed = true;
    return KW_FD1D_OK;
}

// a copy of the options with the exercise flag cleared (src/Pricer/kwFd1d_BlackScholes.cpp:21-28)
__global__ void european_copy_kernel(const kw_option* in, kw_option* out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    kw_option o = load_option(in + i);
    o.e = 0;
    out[i] = o;
}

// options already on the device (`d_opts_p`, copied from `assets`) -> device prices in `d_out` on the handle's
// stream; no final sync.  `european`: host-side compression must group by the key with e = 0.
int price_resident(kw_fd1d_handle* h, const kw_option* assets, size_t n, const kw_option* d_opts_p, double* d_out,
                   double* d_out_eu = nullptr, bool european = false);

// host assets -> device prices in `d_out` (n doubles) on the handle's stream; no final sync
int price_to_device(kw_fd1d_handle* h, const kw_option* assets, size_t n, DevBuf<kw_option>& d_opts,
                    double* d_out, double* d_out_eu = nullptr)
{
    KW_CUDA(h, d_opts.reserve(n));
    KW_CUDA(h, cudaMemcpyAsync(d_opts.p, assets, n * sizeof(kw_option), cudaMemcpyHostToDevice, h->stream));
    return price_resident(h, assets, n, d_opts.p, d_out, d_out_eu);
}

int price_resident(kw_fd1d_handle* h, const kw_option* assets, size_t n, const kw_option* d_opts_p, double* d_out,
                   double* d_out_eu, bool european)
{
    KW_CUDA(h, h->d_status.reserve(16));
    Fd1dBatch B;
    memset(&B, 0, sizeof B);
    B.opts = d_opts_p;
    B.prices = d_out;
    B.prices_eu = d_out_eu;  // non-null: the fused FD1D-BS march (both solutions of every chain)
    B.n_opt = (uint32_t)n;
    B.status = h->d_status.p;
    B.tDim = (int32_t)h->cfg.t_grid_size;
    B.xDim = (int32_t)h->cfg.x_grid_size;
    B.density = h->cfg.density;
    B.scale = h->cfg.scale;
    B.max_mode = h->cfg.exact == 0 ? 4 : (h->cfg.exact == 1 ? 1 : 0);
    h->dev_compressed = false;
    if (h->cfg.compress == 1 && h->layout == KW_FD1D_LAYOUT_REG) {
        if (int rc = compress_on_device(h, B, d_opts_p, n, h->stream)) return rc;
    } else if (h->cfg.compress) {
        size_t m;
        uint32_t *rep, *start, *csr;
        std::vector<kw_option> euro;
        if (european) {  // host-side grouping (FD1D.GPU.COMPRESS = 2 / the SoA layout) of the European copies
            euro.assign(assets, assets + n);
            for (auto& o : euro) o.e = 0;
        }
        if (int rc = compress(h, european ? euro.data() : assets, n, m, rep, start, csr)) return rc;
        if (m < n) {
            KW_CUDA(h, h->d_rep.reserve(m));
            KW_CUDA(h, h->d_start.reserve(m + 1));
            KW_CUDA(h, h->d_csr.reserve(n));
            KW_CUDA(h, cudaMemcpyAsync(h->d_rep.p, rep, m * 4, cudaMemcpyHostToDevice, h->stream));
            KW_CUDA(h, cudaMemcpyAsync(h->d_start.p, start, (m + 1) * 4, cudaMemcpyHostToDevice, h->stream));
            KW_CUDA(h, cudaMemcpyAsync(h->d_csr.p, csr, n * 4, cudaMemcpyHostToDevice, h->stream));
            B.pde_rep = h->d_rep.p;
            B.csr_start = h->d_start.p;
            B.csr_opt = h->d_csr.p;
        }
        B.n_pde = (uint32_t)m;
    } else {
        B.n_pde = (uint32_t)n;
    }
    return launch_batch(h, B, h->stream);
}

int check_status(kw_fd1d_handle* h, cudaStream_t st, const kw_option* host_assets)
{
    KW_CUDA(h, h->h_status.reserve(16));
    KW_CUDA(h, cudaMemcpyAsync(h->h_status.p, h->d_status.p, 10 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    KW_CUDA(h, cudaStreamSynchronize(st));
    h->unsynced = false;
    for (int i = 0; i < 6; ++i) h->mode_count[i] = h->h_status.p[2 + i];
    h->long_chains = h->h_status.p[9];
    if (h->dev_compressed) {
        // the PDE count stayed on the device; it is the sum of the per-mode counters of the march
        uint64_t m = 0;
        for (int i = 0; i < 6; ++i) m += h->mode_count[i];
        h->last_n_pde = m;
        if (h->class_split) h->last_var = m < h->class_split ? h->var_small : h->var;
    }
    if (h->h_status.p[0] != 0) {
        const unsigned int idx = h->h_status.p[1];
        if (host_assets) return fail(h, KW_FD1D_ERANGE, range_message(h, host_assets[idx]));
        return fail(h, KW_FD1D_ERANGE,
                    "Fd1d_Pricer::price Fd1d::value: x not in range (option " + std::to_string(idx) + ")");
    }
    return KW_FD1D_OK;
}


// occupancy and attributes of one kernel variant
int prepare_variant(kw_fd1d_handle* h, const RegVariant* var, const cudaDeviceProp& prop, int& ctas_per_sm, int& regs)
{
    if (var->wide_nwp) {
        KW_CUDA(h, cudaFuncSetAttribute(var->setup_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)var->setup_smem));
        KW_CUDA(h, cudaFuncSetAttribute(var->wide_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)var->smem));
        KW_CUDA(h, cudaFuncSetAttribute(var->wide_fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
        cudaFuncAttributes fa;
        KW_CUDA(h, cudaFuncGetAttributes(&fa, var->wide_fn));
        regs = fa.numRegs;
        const int by_regs = 65536 / (fa.numRegs * 128);
        const int by_smem = (int)((size_t)prop.sharedMemPerMultiprocessor / (var->smem + 1024));
        ctas_per_sm = std::max(1, std::min({by_regs, by_smem, 512 / var->tmem_cols, var->minb}));
        return KW_FD1D_OK;
    }
    KW_CUDA(h, cudaFuncSetAttribute(var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)var->smem));
    cudaFuncAttributes fa;
    KW_CUDA(h, cudaFuncGetAttributes(&fa, var->fn));
    regs = fa.numRegs;
    int occ = 0;
    const int block = var->pdes_per_cta > 1 ? std::max(128, var->P) : var->P;  // Layout W: four warps = four PDEs (eight: fused FD1D-BS)
    KW_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var->fn, block, var->smem));
    if (occ < 1) return fail(h, KW_FD1D_ECUDA, "Fd1dGpu_Pricer::init: kernel variant does not fit on an SM");
    if (var->tmem_cols > 0) {
        // The occupancy calculator answers 1 for any kernel that allocates tensor memory, but the
        // hardware co-schedules CTAs as long as their tcgen05.alloc requests fit the SM's 512 columns
        // (measured: kw_fd1d_tmem_probe finds 4 x 128 columns resident).  Size the persistent grid
        // from the real limits: registers, shared memory, TMEM columns.
        const int by_regs = 65536 / (fa.numRegs * block);
        const int by_smem = (int)((size_t)prop.sharedMemPerMultiprocessor / (var->smem + 1024));
        const int by_tmem = 512 / var->tmem_cols;
        occ = std::max(1, std::min({by_regs, by_smem, by_tmem, var->minb}));
        // ... and ask for the shared-memory carve-out that many CTAs need (the driver would size it
        // for the single CTA the calculator believes in)
        KW_CUDA(h, cudaFuncSetAttribute(var->fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    }
    ctas_per_sm = occ;
    return KW_FD1D_OK;
}

// the first CTA-per-PDE entry of the smallest fitting tile
const RegVariant* find_small_variant(int xDim, int prec)
{
    const RegVariant* best = nullptr;
    for (int i = 0; i < kNumVariants; ++i) {
        const RegVariant& v = g_variants[i];
        if (v.M * v.P < xDim || v.prec != prec || v.pdes_per_cta > 1 || v.tmem_cols > 0 || v.wide_nwp) continue;
        if (!best || v.P < best->P) best = &v;
    }
    return best;
}

}  // namespace

extern "C" {

void kw_fd1d_config_default(kw_fd1d_config* cfg)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->density = 0.25;  // src/Pricer/kwFd1d.cpp:12
    cfg->scale = 50.;     // :13
    cfg->t_grid_size = 512;
    cfg->x_grid_size = 512;
    cfg->device = 0;
    cfg->precision = KW_FD1D_F64;
    cfg->layout = KW_FD1D_LAYOUT_AUTO;
    cfg->compress = 1;
    cfg->variant = 0;
    cfg->exact = 0;
}

int kw_fd1d_create_multi(const kw_fd1d_config* cfg, const int32_t* devices, int32_t n_devices, kw_fd1d_handle** out)
{
    if (!cfg || !out) return KW_FD1D_EINVAL;
    *out = nullptr;
    kw_fd1d_handle* h = new kw_fd1d_handle();
    h->cfg = *cfg;
    *out = h;
    if (!devices || n_devices < 1 || n_devices > 64)
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: FD1D.GPU.DEVICES needs 1 ... 64 device ordinals");
    for (int32_t g = 0; g < n_devices; ++g) {
        kw_fd1d_config c = *cfg;
        c.device = devices[g];
        c.n_devices = 1;
        kw_fd1d_handle* sh = nullptr;
        const int rc = kw_fd1d_create(&c, &sh);
        if (rc != KW_FD1D_OK) {
            const std::string msg = sh ? sh->err : std::string("Fd1dGpu_Pricer::init: out of memory");
            if (sh) kw_fd1d_destroy(sh);
            return fail(h, rc, msg + " (device " + std::to_string(devices[g]) + ")");
        }
        h->shards.push_back(sh);
    }
    h->cfg.device = devices[0];
    h->cfg.n_devices = n_devices;
    return KW_FD1D_OK;
}

int kw_fd1d_create(const kw_fd1d_config* cfg, kw_fd1d_handle** out)
{
    if (!cfg || !out) return KW_FD1D_EINVAL;
    *out = nullptr;
    if (cfg->n_devices > 1 || cfg->n_devices < 0) {
        // cfg.n_devices = N: devices cfg.device ... cfg.device + N - 1; -1: every visible device from cfg.device on
        int ndev = 0;
        int32_t want = cfg->n_devices;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && want < 0) want = std::max(1, ndev - cfg->device);
        if (want > 1 || cfg->n_devices > 1) {
            std::vector<int32_t> devs;
            for (int32_t g = 0; g < std::max(want, 1); ++g) devs.push_back(cfg->device + g);
            return kw_fd1d_create_multi(cfg, devs.data(), (int32_t)devs.size(), out);
        }
    }
    kw_fd1d_handle* h = new kw_fd1d_handle();
    h->cfg = *cfg;
    *out = h;  // returned even on failure so the caller can read the message; destroy it either way
    if (cfg->t_grid_size < 2 || cfg->x_grid_size < 3 || cfg->t_grid_size > (1 << 24) || cfg->x_grid_size > (1 << 20))
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: FD1D.T_GRID_SIZE must be >= 2 and FD1D.X_GRID_SIZE >= 3");
    if (!(cfg->density > 0) || !(cfg->scale > 0))
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: FD1D.DENSITY and FD1D.SCALE must be positive");
    if (cfg->bs_fused < 0 || cfg->bs_fused > 4)
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: FD1D.GPU.BS_FUSED must be 0 ... 4");
    if (cfg->exact < 0 || cfg->exact > 2)
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: FD1D.GPU.EXACT must be 0, 1 or 2");
    if (cfg->precision != KW_FD1D_F64 && cfg->precision != KW_FD1D_F32)
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: FD1D.GPU.PRECISION must be f64 or f32");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(h, KW_FD1D_ECUDA,
                    std::string("Fd1dGpu_Pricer::init: no CUDA device (") + cudaGetErrorString(e) +
                        "); this pricer has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: FD1D.GPU.DEVICE out of range");
    KW_CUDA(h, cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    KW_CUDA(h, cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    strncpy(h->name, prop.name, sizeof h->name - 1);
    KW_CUDA(h, cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, cfg->device));
    if (prop.major < 10)
        return fail(h, KW_FD1D_ECUDA, "Fd1dGpu_Pricer::init: kernels are built for sm_100a only; found " + std::string(prop.name));

    // layout resolution (DESIGN.md "dispatch"): register layout wherever a tile fits
    int layout = cfg->layout;
    if (layout == KW_FD1D_LAYOUT_AUTO) layout = cfg->x_grid_size <= kMaxRegX ? KW_FD1D_LAYOUT_REG : KW_FD1D_LAYOUT_SOA;
    if (layout == KW_FD1D_LAYOUT_REG) {
        h->var = find_variant((int)cfg->x_grid_size, cfg->variant, cfg->precision);
        if (!h->var)
            return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: no register-layout kernel variant for this FD1D.X_GRID_SIZE / FD1D.GPU.VARIANT");
        if (int rc = prepare_variant(h, h->var, prop, h->ctas_per_sm, h->regs)) return rc;
        if (cfg->variant == 0 && (h->var->pdes_per_cta > 1 || h->var->wide_nwp)) {
            // Layout W packs 4 PDEs into a CTA and 8 into an SM: below one full wave of the device the
            // CTA-per-PDE kernel spreads the batch over more SMs (DESIGN.md "dispatch")
            h->var_small = find_small_variant((int)cfg->x_grid_size, cfg->precision);
            if (h->var_small) {
                if (int rc = prepare_variant(h, h->var_small, prop, h->ctas_per_sm_small, h->regs_small)) return rc;
                h->small_below = (uint32_t)(h->sm_count * h->ctas_per_sm * std::max(1, h->var->pdes_per_cta));
                // Layout W for x <= 1024 finishes any batch up to one wave in one group period (1.0 ms at 1024^2,
                // 0.27 ms at 512^2), the CTA-per-PDE kernel needs a second and third round from 444 / 888 PDEs on:
                // measured crossover at ~0.7 of a wave (profiles/r1_bv_small_batch_crossover.log)
                if (h->var->pdes_per_cta == 4 && !h->var->wide_nwp) h->small_below = h->small_below * 3 / 4;
                // packed warps (two / four PDEs per warp): a batch of up to half a wave takes one lone-warp period whatever its
                // size; the CTA-per-PDE kernel wins only while it needs a single round of its own
                if (h->var->pdes_per_cta > 4) h->small_below = (uint32_t)(h->sm_count * h->ctas_per_sm_small) + 1;
            }
        }
        // fused FD1D-BS march: fp64, one Layout W tile of 4 chunks per lane
        const bool tile4 = cfg->x_grid_size > 512 && cfg->x_grid_size <= 1024;
        const bool tile2 = cfg->x_grid_size > 256 && cfg->x_grid_size <= 512;
        const bool tile1 = cfg->x_grid_size <= 256;
        const bool wide2 = cfg->x_grid_size > 1024 && cfg->x_grid_size <= 2048;
        const bool wide4 = cfg->x_grid_size > 2048 && cfg->x_grid_size <= 4096;
        const bool seq = cfg->bs_fused == 0 || cfg->bs_fused == 4;  // march as given, then the European copy
        const bool f32 = cfg->precision == KW_FD1D_F32;
        if (f32 && cfg->bs_fused != 1 && seq && (tile4 || tile2 || tile1) && cfg->variant == 0) {
            h->var_bs = &g_bsf_variant[tile4 ? 0 : (tile2 ? 1 : 2)];
            h->bs_forced = cfg->bs_fused >= 2;
            if (int rc = prepare_variant(h, h->var_bs, prop, h->ctas_per_sm_bs, h->regs_bs)) return rc;
        } else if (cfg->bs_fused != 1 && cfg->precision == KW_FD1D_F64 && (tile4 || ((tile1 || tile2 || wide2 || wide4) && seq)) &&
            (cfg->variant == 0 || cfg->bs_fused >= 2)) {
#ifdef KW_EXPERIMENTS
            h->var_bs = cfg->bs_fused == 2 ? &g_bs1_variant
                                           : (cfg->bs_fused == 3 ? &g_bs_variant
                                              : (tile4 ? &g_bs2_variant : (tile2 ? &g_bs2n2_variant : (tile1 ? &g_bs2n1_variant : (wide2 ? &g_bs_wide2_variant : &g_bs_wide4_variant)))));
#else
            if (cfg->bs_fused == 2 || cfg->bs_fused == 3)
                return fail(h, KW_FD1D_EINVAL,
                            "Fd1dGpu_Pricer::init: FD1D.GPU.BS_FUSED = 2 / 3 are experiments (build with -DKW_EXPERIMENTS)");
            h->var_bs = tile4 ? &g_bs2_variant : (tile2 ? &g_bs2n2_variant : (tile1 ? &g_bs2n1_variant : (wide2 ? &g_bs_wide2_variant : &g_bs_wide4_variant)));
#endif
            h->bs_forced = cfg->bs_fused >= 2;
            if (int rc = prepare_variant(h, h->var_bs, prop, h->ctas_per_sm_bs, h->regs_bs)) return rc;
        }
    } else if (layout == KW_FD1D_LAYOUT_SOA) {
        if (cfg->precision != KW_FD1D_F64)
            return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: the SoA layout (FD1D.GPU.LAYOUT = soa) is fp64 only");
        cudaFuncAttributes fa;
        KW_CUDA(h, cudaFuncGetAttributes(&fa, fd1d_soa_march_kernel));
        h->regs = fa.numRegs;
        h->ctas_per_sm = 0;
    } else {
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::init: unknown FD1D.GPU.LAYOUT");
    }
    h->layout = layout;
    KW_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    KW_CUDA(h, cudaEventCreate(&h->ev0));
    KW_CUDA(h, cudaEventCreate(&h->ev1));
    KW_CUDA(h, h->d_status.reserve(16));
    KW_CUDA(h, h->h_status.reserve(8));
    return KW_FD1D_OK;
}

void kw_fd1d_destroy(kw_fd1d_handle* h)
{
    if (!h) return;
    for (kw_fd1d_handle* sh : h->shards) kw_fd1d_destroy(sh);
    h->shards.clear();
    if (h->stream) {
        cudaSetDevice(h->cfg.device);
        cudaStreamSynchronize(h->stream);
    }
    h->d_opts.release();
    h->d_opts2.release();
    h->d_prices.release();
    h->d_prices2.release();
    h->d_rep.release();
    h->d_start.release();
    h->d_csr.release();
    h->d_chain.release();
    h->d_long.release();
    h->d_long_meta.release();
    h->d_ws.release();
    h->d_status.release();
    h->d_soa.release();
    h->h_idx.release();
    h->h_status.release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// Multi-device price call: contiguous blocks of the batch, one per device, every device driven by its own host
// thread through its single-device shard handle (H2D of the block, chain compression, march, D2H of the block's
// prices straight into the caller's array).  No device-to-device traffic: a PDE never reads another PDE's data
// (reference src/Math/kwFd1d.cpp:61-136), so the only "gather" is that every shard writes its slice of `prices`.
// Small batches use fewer devices: every device in use gets at least two waves of its persistent grid, so the
// kernel each block runs is the one the whole batch would run on one device (bit-identical prices).
static int price_multi(kw_fd1d_handle* h, const kw_option* assets, size_t n, double* prices, bool bs)
{
    h->err.clear();
    h->launches = 0;
    if (n == 0) return KW_FD1D_OK;
    if (!assets || !prices) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: null buffer");
    if (n > 0xfffffff0ull) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: batch too large");
    const kw_fd1d_handle* s0 = h->shards[0];
    const size_t wave = (size_t)std::max(1, s0->sm_count) * std::max(1, s0->ctas_per_sm) *
                        (s0->var ? std::max(1, s0->var->pdes_per_cta) : 1);
    size_t use = std::min<size_t>(h->shards.size(), std::max<size_t>(1, n / (2 * wave)));
    if (h->cfg.variant != 0) use = std::min(h->shards.size(), n);  // pinned kernel: any split gives the same bits
    h->shards_used = (uint32_t)use;
    std::vector<int> rc(use, KW_FD1D_OK);
    auto run = [&](size_t g) {
        const size_t lo = n * g / use, hi = n * (g + 1) / use;
        kw_fd1d_handle* sh = h->shards[g];
        rc[g] = bs ? kw_fd1d_price_bs(sh, assets + lo, hi - lo, prices + lo) : kw_fd1d_price(sh, assets + lo, hi - lo, prices + lo);
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (size_t g = 1; g < use; ++g) th.emplace_back(run, g);
    run(0);
    for (auto& t : th) t.join();
    h->last_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    h->last_n_pde = 0;
    for (size_t g = 0; g < use; ++g) {
        h->launches += h->shards[g]->launches;
        h->last_n_pde += h->shards[g]->last_n_pde;
    }
    for (int i = 0; i < 6; ++i) {
        h->mode_count[i] = 0;
        for (size_t g = 0; g < use; ++g) h->mode_count[i] += h->shards[g]->mode_count[i];
    }
    // the reference reports the FIRST failing option (src/Pricer/kwFd1d.cpp:154-155): blocks are contiguous and in
    // order, so that is the error of the lowest shard that has one
    for (size_t g = 0; g < use; ++g)
        if (rc[g] != KW_FD1D_OK) return fail(h, rc[g], h->shards[g]->err);
    return KW_FD1D_OK;
}

int kw_fd1d_price(kw_fd1d_handle* h, const kw_option* assets, size_t n, double* prices)
{
    if (h && !h->shards.empty()) return price_multi(h, assets, n, prices, false);
    if (h) h->launches = 0;
    if (!h) return KW_FD1D_EINVAL;
    if (!h->stream) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: pricer was not initialised");
    h->err.clear();
    if (n == 0) return KW_FD1D_OK;  // src/Pricer/kwFd1d.cpp:24-26
    if (!assets || !prices) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: null buffer");
    if (n > 0xfffffff0ull) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: batch too large");
    KW_CUDA(h, cudaSetDevice(h->cfg.device));
    KW_CUDA(h, h->d_prices.reserve(n));
    if (int rc = price_to_device(h, assets, n, h->d_opts, h->d_prices.p)) return rc;
    KW_CUDA(h, cudaMemcpyAsync(prices, h->d_prices.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return check_status(h, h->stream, assets);
}

int kw_fd1d_price_bs(kw_fd1d_handle* h, const kw_option* assets, size_t n, double* prices)
{
    if (h && !h->shards.empty()) return price_multi(h, assets, n, prices, true);
    if (h) h->launches = 0;
    if (!h) return KW_FD1D_EINVAL;
    if (!h->stream) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: pricer was not initialised");
    h->err.clear();
    if (n == 0) return KW_FD1D_OK;
    if (!assets || !prices) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: null buffer");
    if (n > 0xfffffff0ull) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: batch too large");
    KW_CUDA(h, cudaSetDevice(h->cfg.device));
    KW_CUDA(h, h->d_prices.reserve(n));
    KW_CUDA(h, h->d_prices2.reserve(n));
    if (h->cfg.bs_fused >= 2 && !h->var_bs)
        return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: FD1D.GPU.BS_FUSED = 2 / 3 needs fp64 and 512 < FD1D.X_GRID_SIZE <= 1024, = 4 fp64 and FD1D.X_GRID_SIZE <= 4096");
    // auto: fused from one full wave of the persistent grid (4 chains per CTA) upwards; below that the
    // CTA-per-PDE kernel of the two-solve path spreads the batch over more SMs
    if (h->var_bs && (h->bs_forced || 4 * n >= (size_t)h->sm_count * h->ctas_per_sm_bs * std::min(4, h->var_bs->pdes_per_cta) * 3)) {  // 3/4 of a wave, as small_below
        // fused: the solve as given (:18) and the solve of the European copies (:21-28) are two value
        // vectors of the same chains marched by one launch; then + (BS - FD_euro) (:30-40)
        if (int rc = price_to_device(h, assets, n, h->d_opts, h->d_prices.p, h->d_prices2.p)) return rc;
        bs_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_opts.p, n, h->d_prices.p, h->d_prices2.p);
        h->launches += 1;
        KW_CUDA(h, cudaGetLastError());
        KW_CUDA(h, cudaMemcpyAsync(prices, h->d_prices.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        return check_status(h, h->stream, assets);
    }
    // 1. FD as given (src/Pricer/kwFd1d_BlackScholes.cpp:18)
    if (int rc = price_to_device(h, assets, n, h->d_opts, h->d_prices.p)) return rc;
    if (int rc = check_status(h, h->stream, assets)) {
        // the header's promise: the other prices are still written (the failing ones as NaN)
        if (rc == KW_FD1D_ERANGE) {
            const std::string msg = h->err;
            cudaMemcpyAsync(prices, h->d_prices.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
            cudaStreamSynchronize(h->stream);
            h->err = msg;
        }
        return rc;
    }
    // 2. FD on European copies (:21-28): the copy is made on the device from the options already there
    KW_CUDA(h, h->d_opts2.reserve(n));
    european_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_opts.p, h->d_opts2.p, n);
    h->launches += 1;
    if (int rc = price_resident(h, assets, n, h->d_opts2.p, h->d_prices2.p, nullptr, /*european=*/true)) return rc;
    // 3. + (BS - FD_euro) (:30-40)
    bs_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_opts2.p, n, h->d_prices.p, h->d_prices2.p);
    h->launches += 1;
    KW_CUDA(h, cudaGetLastError());
    KW_CUDA(h, cudaMemcpyAsync(prices, h->d_prices.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return check_status(h, h->stream, assets);
}

int kw_fd1d_price_device(kw_fd1d_handle* h, const kw_option* d_assets, size_t n, double* d_prices, void* stream)
{
    if (h) h->launches = 0;
    if (!h) return KW_FD1D_EINVAL;
    if (!h->shards.empty())
        return fail(h, KW_FD1D_EINVAL,
                    "Fd1dGpu_Pricer::price: device-resident buffers live on one device; use a single-device handle per GPU");
    if (n > 0xfffffff0ull) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: batch too large");
    if (!h->stream) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: pricer was not initialised");
    h->err.clear();
    if (n == 0) return KW_FD1D_OK;
    if (!d_assets || !d_prices) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer::price: null buffer");
    KW_CUDA(h, cudaSetDevice(h->cfg.device));
    Fd1dBatch B;
    memset(&B, 0, sizeof B);
    B.opts = d_assets;
    B.prices = d_prices;
    B.status = h->d_status.p;
    B.n_opt = (uint32_t)n;
    B.n_pde = (uint32_t)n;
    B.tDim = (int32_t)h->cfg.t_grid_size;
    B.xDim = (int32_t)h->cfg.x_grid_size;
    B.density = h->cfg.density;
    B.scale = h->cfg.scale;
    B.max_mode = h->cfg.exact == 0 ? 4 : (h->cfg.exact == 1 ? 1 : 0);
    h->dev_compressed = false;
    if (h->cfg.compress == 1 && h->layout == KW_FD1D_LAYOUT_REG)
        if (int rc = compress_on_device(h, B, d_assets, n, (cudaStream_t)stream)) return rc;
    return launch_batch(h, B, (cudaStream_t)stream);
}

int kw_fd1d_sync(kw_fd1d_handle* h, void* stream)
{
    if (!h) return KW_FD1D_EINVAL;
    if (!h->shards.empty()) return KW_FD1D_OK;  // multi-device price calls return synchronised
    if (!h->stream) return fail(h, KW_FD1D_EINVAL, "Fd1dGpu_Pricer: pricer was not initialised");
    KW_CUDA(h, cudaSetDevice(h->cfg.device));
    return check_status(h, (cudaStream_t)stream, nullptr);
}

int kw_fd1d_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int kw_fd1d_get_device_props(int32_t device, kw_fd1d_device_props* out)
{
    if (!out) return KW_FD1D_EINVAL;
    memset(out, 0, sizeof *out);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return KW_FD1D_ECUDA;
    strncpy(out->name, p.name, sizeof out->name - 1);
    out->integrated = p.integrated;
    out->sm_count = p.multiProcessorCount;
    out->regs_per_sm = p.regsPerMultiprocessor;
    out->max_blocks_per_sm = p.maxBlocksPerMultiProcessor;
    out->max_threads_per_sm = p.maxThreadsPerMultiProcessor;
    out->mem_bus_width_bits = p.memoryBusWidth;
    out->total_mem_bytes = p.totalGlobalMem;
    cudaDeviceGetAttribute(&out->clock_khz, cudaDevAttrClockRate, device);
    cudaDeviceGetAttribute(&out->mem_clock_khz, cudaDevAttrMemoryClockRate, device);
    return KW_FD1D_OK;
}

const char* kw_fd1d_last_error(const kw_fd1d_handle* h) { return h ? h->err.c_str() : "null handle"; }

int kw_fd1d_get_info(const kw_fd1d_handle* hc, kw_fd1d_info* info)
{
    if (!hc || !info) return KW_FD1D_EINVAL;
    kw_fd1d_handle* h = const_cast<kw_fd1d_handle*>(hc);
    if (!h->shards.empty()) {
        // multi-device: the first shard's kernel facts, totals over the shards the last call used
        if (int rc = kw_fd1d_get_info(h->shards[0], info)) return rc;
        const uint32_t used = std::max<uint32_t>(1, h->shards_used);
        for (uint32_t g = 1; g < used; ++g) {
            kw_fd1d_info gi;
            if (kw_fd1d_get_info(h->shards[g], &gi) == KW_FD1D_OK) info->last_kernel_ms = std::max(info->last_kernel_ms, gi.last_kernel_ms);
        }
        info->launches = h->launches;
        info->last_n_pde = h->last_n_pde;
        for (int i = 0; i < 6; ++i) info->mode_count[i] = h->mode_count[i];
        info->n_devices = (int32_t)h->shards.size();
        info->devices_used = (int32_t)used;
        info->last_wall_ms = h->last_wall_ms;
        for (uint32_t g = 1; g < used; ++g) info->long_chains += h->shards[g]->long_chains;
        return KW_FD1D_OK;
    }
    memset(info, 0, sizeof *info);
    info->n_devices = 1;
    info->devices_used = 1;
    info->device = h->cfg.device;
    info->sm_count = h->sm_count;
    info->layout = h->layout;
    const RegVariant* lv = h->last_var ? h->last_var : h->var;
    const bool lsmall = lv && lv == h->var_small;
    info->variant = lv ? lv->id : 0;
    const bool lw = lv && (lv->pdes_per_cta > 1 || lv->wide_nwp);  // Layout W: 32 nodes per lane
    info->threads_per_pde = lv ? (lv->wide_nwp ? 32 * lv->wide_nwp : (lw ? 128 / lv->pdes_per_cta : lv->P)) : 1;
    info->nodes_per_thread = lv ? (lw ? 32 : lv->M) : (int)h->cfg.x_grid_size;
    const bool lbs = lv && lv == h->var_bs;
    info->ctas_per_sm = lbs ? h->ctas_per_sm_bs : (lsmall ? h->ctas_per_sm_small : h->ctas_per_sm);
    info->regs_per_thread = lbs ? h->regs_bs : (lsmall ? h->regs_small : h->regs);
    info->smem_per_cta = lv ? (int)lv->smem : 0;
    info->grid = h->last_grid;
    info->sm_clock_khz = h->clock_khz;
    info->launches = h->launches;
    info->last_n_pde = h->last_n_pde;
    for (int i = 0; i < 6; ++i) info->mode_count[i] = h->mode_count[i];
    info->long_chains = h->long_chains;
    info->last_kernel_ms = 0.;
    if (h->ev_valid && cudaEventSynchronize(h->ev1) == cudaSuccess) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) info->last_kernel_ms = ms;
    }
    strncpy(info->device_name, h->name, sizeof info->device_name - 1);
    return KW_FD1D_OK;
}

int kw_fd1d_fp64_peak(int32_t device, double* tflops, double* sm_mhz_effective)
{
    if (!tflops) return KW_FD1D_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return KW_FD1D_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return KW_FD1D_ECUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return KW_FD1D_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = prop.multiProcessorCount * 2, tpb = 1024, iters = 4096;
    dfma_throughput_kernel<<<grid, tpb>>>(d, 64, 1.0);  // warm-up
    double best = 0.;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        dfma_throughput_kernel<<<grid, tpb>>>(d, iters, 1.0);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 64.0 * (double)iters * (double)grid * tpb;
        best = std::max(best, flop / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || best == 0.) return KW_FD1D_ECUDA;
    *tflops = best;
    if (sm_mhz_effective) *sm_mhz_effective = best * 1e12 / (2.0 * 64.0 * prop.multiProcessorCount) * 1e-6;
    return KW_FD1D_OK;
}

int kw_fd1d_microbench(int32_t device, double* out8)
{
    if (!out8) return KW_FD1D_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return KW_FD1D_ECUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, 8 * sizeof(double)) != cudaSuccess) return KW_FD1D_ECUDA;
    cudaMemset(d, 0, 8 * sizeof(double));
    latency_kernel<<<1, 128>>>(d, 1.0, 8);
    latency_kernel<<<1, 128>>>(d, 1.0, 256);
    cudaError_t e = cudaMemcpy(out8, d, 8 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? KW_FD1D_OK : KW_FD1D_ECUDA;
}


int kw_fd1d_tmem_probe(int32_t device, double* out16)
{
    if (!out16) return KW_FD1D_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return KW_FD1D_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return KW_FD1D_ECUDA;
    const int grid = prop.multiProcessorCount * 4;
    const size_t nrec = 3 + 3 * (size_t)grid;
    double* d = nullptr;
    long long* cyc = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess || cudaMalloc(&cyc, nrec * 8) != cudaSuccess) return KW_FD1D_ECUDA;
    typedef void (*Probe)(double*, long long*, int, double);
    const Probe probes[6] = {tmem_probe_kernel<0>, tmem_probe_kernel<1>, tmem_probe_kernel<2>,
                             tmem_probe_kernel<3>, tmem_probe_kernel<4>, tmem_probe_kernel<5>};
    const int iters = 2000;
    double mism = 0.;
    std::vector<long long> hc(nrec);
    for (int i = 0; i < 16; ++i) out16[i] = 0.;
    for (int m = 0; m < 7; ++m) {
        cudaMemset(cyc, 0, nrec * 8);
        if (m < 6)
            probes[m]<<<grid, 128>>>(d, cyc, iters, 1.0);
        else
            probes[1]<<<1, 32>>>(d, cyc, iters, 1.0);  // one warp alone: ld + wait round trip
        if (cudaMemcpy(hc.data(), cyc, nrec * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaFree(d);
            cudaFree(cyc);
            return KW_FD1D_ECUDA;
        }
        out16[m] = (double)hc[0] / iters;  // cycles per round of 8 chunks
        mism += (double)hc[1];
        if (m < 6) {
            // how many CTAs shared CTA 0's SM for at least half of its timed loop
            const long long sm0 = hc[2], a0 = hc[3], a1 = hc[4];
            int co = 0;
            for (int b = 0; b < grid; ++b) {
                if (hc[2 + 3 * b] != sm0) continue;
                const long long lo = std::max(a0, hc[3 + 3 * b]), hi = std::min(a1, hc[4 + 3 * b]);
                if (hi - lo > (a1 - a0) / 2) ++co;
            }
            out16[8 + m] = co;
            if (getenv("KW_PROBE_DEBUG")) {
                for (int b = 0; b < grid; ++b)
                    if (hc[2 + 3 * b] == sm0)
                        fprintf(stderr, "probe %d: sm %lld cta %d start %+lld dur %lld\n", m, sm0, b,
                                hc[3 + 3 * b] - a0, hc[4 + 3 * b] - hc[3 + 3 * b]);
            }
            if (m == 0 && hc[2 + 3 * grid] > 0) out16[15] = (double)hc[0] / (double)hc[2 + 3 * grid];  // clock64 ticks per ns
        }
    }
    out16[7] = mism;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tmem_probe_kernel<3>, 128, 0);
    out16[14] = occ;
    cudaFree(d);
    cudaFree(cyc);
    return KW_FD1D_OK;
}

int kw_fd1d_dfma_probe(int32_t device, double* out8)
{
    if (!out8) return KW_FD1D_EINVAL;
    for (int i = 0; i < 10; ++i) out8[i] = 0.;
    if (cudaSetDevice(device) != cudaSuccess) return KW_FD1D_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return KW_FD1D_ECUDA;
    double *din = nullptr, *dout = nullptr;
    if (cudaMalloc(&din, 64 * 8) != cudaSuccess || cudaMalloc(&dout, 64) != cudaSuccess) return KW_FD1D_ECUDA;
    double hin[64];
    for (int i = 0; i < 64; ++i) hin[i] = 1e-3 * (i + 1);
    cudaMemcpy(din, hin, sizeof hin, cudaMemcpyHostToDevice);
    typedef void (*K)(const double*, double*, int);
    const K ks[6] = {dfma_operand_kernel<1>, dfma_operand_kernel<2>, dfma_operand_kernel<3>,
                     dfma_operand_kernel<4>, dfma_operand_kernel<5>, dfma_operand_kernel<6>};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096;
    for (int occ = 0; occ < 2; ++occ) {             // 16 and 64 warps per SM
        const int ctas = occ == 0 ? 4 : 16;
        const int grid = prop.multiProcessorCount * ctas;
        for (int k = 0; k < 6; ++k) {
            if (k >= 4 && occ == 1) continue;  // the select probes run at 16 warps per SM only
            ks[k]<<<grid, 128>>>(din, dout, 64);
            cudaEventRecord(e0);
            ks[k]<<<grid, 128>>>(din, dout, iters);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) return KW_FD1D_ECUDA;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            out8[k < 4 ? occ * 4 + k : 4 + k] = 2.0 * 32.0 * iters * (double)grid * 128 / (ms * 1e-3) * 1e-12;  // TFLOP/s
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(din);
    cudaFree(dout);
    return KW_FD1D_OK;
}

int kw_fd1d_has_variant(int32_t id, int32_t precision)
{
    for (int i = 0; i < kNumVariants; ++i)
        if (g_variants[i].id == id && g_variants[i].prec == precision) return 1;
    if (precision == KW_FD1D_F32 && (id == 1257 || id == 1158 || id == 1058)) return 1;
    if (precision == KW_FD1D_F64 && (id == g_bs2_variant.id || id == g_bs2n2_variant.id || id == g_bs2n1_variant.id || id == g_bs_wide2_variant.id || id == g_bs_wide4_variant.id)) return 1;
#ifdef KW_EXPERIMENTS
    if (precision == KW_FD1D_F64 && (id == g_bs_variant.id || id == g_bs1_variant.id || id == g_bs253_variant.id)) return 1;
#endif
    return 0;
}

#ifdef KW_EXPERIMENTS
const char* kw_fd1d_version(void) { return "kwinto-b200 fd1d 0.2 (sm_100a, +experiments)"; }
#else
const char* kw_fd1d_version(void) { return "kwinto-b200 fd1d 0.2 (sm_100a)"; }
#endif

}  // extern "C"
