/*
 * kw_fd1d.h -- C ABI of the B200-native Fd1d (1-D finite-difference) option pricer.
 *
 * This is the drop-in boundary for ONE path of gituliar/kwinto-cuda: the pricer behind
 *   kw::Pricer::init  / kw::Pricer::price            (reference src/Pricer/kwPricer.h:12-22)
 * as implemented for mode "FD1D" by
 *   kw::Fd1d_Pricer::init / ::price                   (reference src/Pricer/kwFd1d.cpp:9-19, :21-160)
 *   kw::Fd1d::solve / ::solveOne / ::value            (reference src/Math/kwFd1d.cpp:11, :61, :139)
 *   kw::solveTridiagonal                              (reference src/Math/kwMath.cpp:16-49)
 * The reference has no FFI of its own (it is one C++ process); the entry points below are
 * what a `class Fd1dGpu_Pricer : public kw::Pricer` registered in PricerFactory::create
 * (reference src/Pricer/kwPricerFactory.h:15-41) binds -- see INTEGRATION.md and
 * kwinto-cuda_b200/host/kw/Pricer/kwFd1dGpu.h for that subclass.
 *
 * Plain pointers and sizes only; no C++ or torch types.  All functions return 0 on success
 * and a non-zero KW_FD1D_E* code otherwise; the message (the reference's `Error` string,
 * src/Core/kwTypes.h:18, "" = success) is retrieved with kw_fd1d_last_error().
 * A handle is NOT thread-safe (neither is the reference pricer: singleton thread pool with a
 * global wait(), src/kwThreadPool.cpp:58-72); use one handle per host thread / per GPU.
 */
#ifndef KW_FD1D_H
#define KW_FD1D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to kw::Option (reference src/Core/kwAsset.h:12-22): 56 bytes, align 8,
 * offsets t0 k8 z16 r24 q32 s40 e48 w49, so std::vector<kw::Option>::data() passes through. */
typedef struct kw_option {
    double t;  /* time to maturity, years          */
    double k;  /* strike                            */
    double z;  /* volatility                        */
    double r;  /* interest rate                     */
    double q;  /* dividend rate                     */
    double s;  /* spot                              */
    uint8_t e; /* early exercise: 0 European, 1 American */
    int8_t w;  /* parity: -1 put, +1 call           */
} kw_option;

enum {
    KW_FD1D_OK = 0,
    KW_FD1D_EINVAL = 1,  /* bad argument / configuration                                  */
    KW_FD1D_ECUDA = 2,   /* CUDA runtime failure (message carries cudaGetErrorString)      */
    KW_FD1D_ERANGE = 3,  /* log(s/k) outside the x grid: the reference's Fd1d::value error */
    KW_FD1D_ENOMEM = 4
};

/* kernel layouts (north_star: two layouts are built and compared) */
enum {
    KW_FD1D_LAYOUT_AUTO = 0, /* per-xDim dispatch table, see DESIGN.md                         */
    KW_FD1D_LAYOUT_REG = 1,  /* "B": CTA per PDE, x-grid + LU coefficients in registers/smem,  */
                             /*      partitioned Thomas with warp-shuffle carries             */
    KW_FD1D_LAYOUT_SOA = 2   /* "A": thread per PDE, batch-interleaved SoA in global memory   */
};

enum {
    KW_FD1D_F64 = 0, /* everything in fp64 (parity bar 1e-9 absolute)                        */
    KW_FD1D_F32 = 1  /* fp64 set-up, fp32 time march (parity bar 1e-4 relative)              */
};

/* Mirrors the keys Fd1d_Pricer::init reads (reference src/Pricer/kwFd1d.cpp:12-16) plus the
 * device-side knobs.  kw_fd1d_config_default() fills the reference's defaults. */
typedef struct kw_fd1d_config {
    double density;      /* FD1D.DENSITY      default 0.25 */
    double scale;        /* FD1D.SCALE        default 50.  */
    int64_t t_grid_size; /* FD1D.T_GRID_SIZE  default 512  */
    int64_t x_grid_size; /* FD1D.X_GRID_SIZE  default 512  */
    int32_t device;      /* CUDA device ordinal, default 0 (FD1D.GPU.DEVICE)                 */
    int32_t precision;   /* KW_FD1D_F64 | KW_FD1D_F32 (FD1D.GPU.PRECISION)                   */
    int32_t layout;      /* KW_FD1D_LAYOUT_* (FD1D.GPU.LAYOUT)                               */
    int32_t compress;    /* 1 (default): one PDE per (t,r,q,z,e,w) chain, as the reference's */
                         /* compression (src/Pricer/kwFd1d.cpp:28-65), grouped ON THE DEVICE  */
                         /* (hash join in HBM; host-side for the SoA layout); 2: grouped on   */
                         /* the host; 0: one PDE per option                                   */
    int32_t variant;     /* 0 = auto; otherwise a kernel variant id for tuning   */
    int32_t exact;       /* FD1D.GPU.EXACT: 0 (default) carry terms proven < 2^-56 may be    */
                         /* dropped (DESIGN.md "Truncation"); 1: keep all scan levels;       */
                         /* 2: keep every carry term                                         */
    int32_t bs_fused;    /* FD1D.GPU.BS_FUSED, kw_fd1d_price_bs only.  0 (default): batches of a device wave  */
                         /* or more (fp64, 512 < x <= 1024) are priced by ONE launch in which every warp     */
                         /* marches its chain as given and then the European copy, with one set-up and one   */
                         /* tensor-memory copy of the coefficients (variant 257); other batches: two solves, */
                         /* as the reference.  1: always two solves.  4: variant 253 for every batch size.   */
                         /* Measured experiments with the same prices: 3 = the two marches side  */
                         /* by side in warps w and w + 4 (variant 252), 2 = both in one warp's step (251)    */
    int32_t n_devices;   /* FD1D.GPU.DEVICES: 0 / 1 = one device (`device`); N > 1 = the devices `device` ... `device` + N - 1;   */
                         /* -1 = every visible device from `device` on.  With more than one device the handle owns one  */
                         /* complete per-device context (stream, device buffers, pinned staging) per GPU and            */
                         /* kw_fd1d_price / kw_fd1d_price_bs split the batch into contiguous blocks, price them          */
                         /* concurrently and land every block's prices in the caller's array (kw_fd1d_create_multi)      */
} kw_fd1d_config;

typedef struct kw_fd1d_handle kw_fd1d_handle;

/* What the handle resolved to (for benchmark reports). */
typedef struct kw_fd1d_info {
    int32_t device;
    int32_t sm_count;
    int32_t layout;           /* resolved layout                                             */
    int32_t variant;          /* resolved kernel variant id                                  */
    int32_t threads_per_pde;  /* layout B: CTA size; layout A: 1                             */
    int32_t nodes_per_thread; /* layout B: M                                                 */
    int32_t ctas_per_sm;      /* occupancy the persistent grid was sized for                 */
    int32_t regs_per_thread;
    int32_t smem_per_cta;     /* bytes                                                       */
    int32_t grid;             /* CTAs launched for the last batch                            */
    int32_t sm_clock_khz;     /* cudaDevAttrClockRate                                        */
    int32_t launches;         /* kernels launched by the last price call                     */
    double last_kernel_ms;    /* device time of the last batch's march launch(es), CUDA events */
    uint64_t last_n_pde;      /* PDEs solved by the last price call                          */
    uint32_t mode_count[6];   /* layout B: PDEs of the last SYNCHRONISED call per carry mode 0..4 */
    char device_name[128];
    int32_t n_devices;        /* devices the handle owns (1 for a single-device handle)      */
    int32_t devices_used;     /* devices the last price call spread the batch over            */
    double last_wall_ms;      /* multi-device: host wall time of the last price call (last_kernel_ms = max over devices) */
    uint32_t long_chains;     /* chains of the last SYNCHRONISED call whose options (> 512) were priced by the CTA-per-chain
                                 value kernel instead of the marching warp (fused FD1D-BS counts both solutions) */
    uint32_t reserved_;
} kw_fd1d_info;

void kw_fd1d_config_default(kw_fd1d_config* cfg);

/* Fd1d_Pricer::init (reference src/Pricer/kwFd1d.cpp:9-19) + device set-up. */
int kw_fd1d_create(const kw_fd1d_config* cfg, kw_fd1d_handle** out);
/* The same for an explicit list of device ordinals (cfg.device / cfg.n_devices are ignored; an ordinal may repeat:
 * two shards on one GPU share it through two streams).  The reference prices a portfolio with ONE call,
 * Portfolio::price -> Pricer::price (reference src/Utils/kwPortfolio.cpp:87-98): with this handle that one call
 * uses every listed GPU -- contiguous blocks of the batch, one host thread, stream and buffer set per device, no
 * device-to-device traffic (a PDE never reads another PDE's data, reference src/Math/kwFd1d.cpp:61-136); the
 * "gather" is each device's D2H into its slice of `prices`.  kw_fd1d_price_device needs a single-device handle. */
int kw_fd1d_create_multi(const kw_fd1d_config* cfg, const int32_t* devices, int32_t n_devices, kw_fd1d_handle** out);
void kw_fd1d_destroy(kw_fd1d_handle* h);

/* Fd1d_Pricer::price (reference src/Pricer/kwFd1d.cpp:21-160).  HOST buffers: `assets` has n
 * options, `prices` room for n doubles (the caller's resized vector).  n == 0 returns OK and
 * writes nothing (:24-26).  If any log(s/k) lies outside its grid the call returns
 * KW_FD1D_ERANGE with the reference's message for the first such option (:154-155); the
 * other prices are still written, the failing ones as NaN.
 * Does H2D of the options, the whole time march on the device, D2H of the prices. */
int kw_fd1d_price(kw_fd1d_handle* h, const kw_option* assets, size_t n, double* prices);

/* Same computation with DEVICE-resident buffers (chain compression on the device when
 * cfg.compress == 1, otherwise one PDE per option), enqueued on `stream` (a cudaStream_t, may be 0) without
 * synchronising.  Range errors are reported by the next kw_fd1d_sync(): several batches may be enqueued before one sync;
 * the count and the smallest failing index (relative to its own batch) accumulate over them.
 * ONE STREAM PER HANDLE: the handle's scratch (chain tables, set-up workspace, status block) is shared by its calls, so
 * all kw_fd1d_price_device calls of a handle must be issued to the same stream (or be ordered by events); for concurrent
 * streams use one handle per stream.  n <= 0xfffffff0. */
int kw_fd1d_price_device(kw_fd1d_handle* h, const kw_option* d_assets, size_t n, double* d_prices,
                         void* stream);
/* Waits for `stream`, then reports a pending range error (KW_FD1D_ERANGE) if any. */
int kw_fd1d_sync(kw_fd1d_handle* h, void* stream);

/* Fd1d_BlackScholes_Pricer::price (reference src/Pricer/kwFd1d_BlackScholes.cpp:15-43), mode
 * "FD1D-BS": FD(as given) + (BS_european - FD_european), Black-Scholes closed form as
 * src/Pricer/kwBlackScholes.cpp:27-50.  HOST buffers. */
int kw_fd1d_price_bs(kw_fd1d_handle* h, const kw_option* assets, size_t n, double* prices);

/* Device facts for the driver's "Devices Info" block (the historic `kwinto bench -v` output,
 * reference log/z800_1024_32768.log:21-36). */
typedef struct kw_fd1d_device_props {
    char name[128];
    int32_t integrated;
    int32_t sm_count;
    int32_t clock_khz;
    int32_t regs_per_sm;
    int32_t max_blocks_per_sm;
    int32_t max_threads_per_sm;
    int32_t mem_clock_khz;
    int32_t mem_bus_width_bits;
    int32_t reserved;
    uint64_t total_mem_bytes;
} kw_fd1d_device_props;
int kw_fd1d_device_count(void);
int kw_fd1d_get_device_props(int32_t device, kw_fd1d_device_props* props);

const char* kw_fd1d_last_error(const kw_fd1d_handle* h);
int kw_fd1d_get_info(const kw_fd1d_handle* h, kw_fd1d_info* info);

/* Measured FP64 FMA throughput of `device` (independent DFMA chains on every SM), in TFLOP/s:
 * the roofline denominator bench.py reports next to the nominal 148 SM x 64 DFMA/clk figure. */
int kw_fd1d_fp64_peak(int32_t device, double* tflops, double* sm_mhz_effective);

/* Latency micro-probes used by DESIGN.md's performance model (cycles): dependent DFMA,
 * 64-bit __shfl_up, __syncthreads with 4 warps, LDS.  out[8]. */
int kw_fd1d_microbench(int32_t device, double* out8);

/* Tensor-memory probes (cycles per round of 64 doubles per thread read back from TMEM, 4 CTAs of
 * 128 threads per SM): [0] DFMA only, [1] tcgen05.ld only, [2] ld + 8 DFMA per 8 doubles,
 * [3] ld + 16 DFMA, [4] as 1, [5] ld issued one chunk ahead + 16 DFMA, [6] one warp alone,
 * [7] read-back mismatches (must be 0), [8..13] CTAs co-resident with CTA 0 in probes 0..5,
 * [14] occupancy the runtime reports.  Used to decide whether TMEM can hold the time-invariant
 * coefficient arrays of the march.  out[16]. */
int kw_fd1d_tmem_probe(int32_t device, double* out16);

/* DFMA throughput (TFLOP/s) against the number of distinct REGISTER source operands:
 * out[0..3] with 16 warps per SM: one register source, two, three distinct, three with one shared by
 * consecutive instructions; out[4..7] the same with 64 warps per SM; out[8], out[9]: three-source DFMAs
 * with one / two unrelated 32-bit selects per DFMA in the same stream (16 warps per SM).  The march's
 * DFMAs all read three registers, so this -- not the one-register figure of kw_fd1d_fp64_peak -- is its
 * practical ceiling.  out[10]. */
int kw_fd1d_dfma_probe(int32_t device, double* out10);

/* 1 if this build of the library carries kernel variant `id` for `precision` (cfg.variant), else 0.  The default
 * build ships the variants the dispatch table can reach; the other variants measured in DESIGN.md are compiled
 * with -DKW_EXPERIMENTS (make -C kwinto-cuda_b200/csrc EXPERIMENTS=1). */
int kw_fd1d_has_variant(int32_t id, int32_t precision);

const char* kw_fd1d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KW_FD1D_H */
