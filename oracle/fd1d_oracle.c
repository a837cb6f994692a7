/*
 * fd1d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's 1-D finite-difference
 * (Crank-Nicolson + explicit early-exercise projection) option pricer.
 * It exists so the CUDA path can be checked for parity on a box where
 * /root/reference is not mounted.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (kwinto-cuda_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit
 * against (a) outputs of the unmodified reference compiled into
 * oracle/_ref/libkwref.so (see oracle/Makefile) whenever that library is
 * present and (b) the committed golden vectors tests/golden/ (npz files) that were
 * produced by that same reference build (tests/golden/make_golden.py), and
 * against the reference's own known-answer tests (test/kwPricer_test.cpp:15-29,
 * tolerance 1.3e-3) and the QuantLib fixture bar (test/kwPortfolio_test.cpp:56).
 *
 * Every function cites the reference file:line it follows.  The evaluation
 * order of every floating-point expression is the reference's, so that with
 * the same libm (glibc) and no FMA contraction (-ffp-contract=off, no
 * -march=native) the prices are bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* Wire struct: layout-identical to kw::Option (src/Core/kwAsset.h:12-22):
 * 56 bytes, offsets t0 k8 z16 r24 q32 s40 e48 w49. */
typedef struct {
    double t, k, z, r, q, s;
    uint8_t e;
    int8_t w;
} kwo_option;

static void set_err(char* err, size_t errlen, const char* msg)
{
    if (err && errlen) {
        strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

/* kw::solveTridiagonal, src/Math/kwMath.cpp:16-49 (NR "tridag").
 * gam is caller-provided scratch of xDim doubles (the reference heap-allocates
 * it per call, :26).  Returns 0, or the reference's error number 1/2/3. */
int kwo_solve_tridiagonal(int xDim, const double* al, const double* a, const double* au,
                          const double* y, double* x, double* gam)
{
    if (a[0] == 0)
        return 1; /* :18-19 */
    if (xDim <= 2)
        return 3; /* :21-23 */

    double bet;
    x[0] = y[0] / (bet = a[0]); /* :29 */
    for (int j = 1; j < xDim; j++) {
        gam[j] = au[j - 1] / bet;      /* :32 */
        bet = a[j] - al[j] * gam[j];   /* :33 */
        if (bet == 0)
            return 2;                  /* :35-36 */
        x[j] = (y[j] - al[j] * x[j - 1]) / bet; /* :38 */
    }
    for (int j = xDim - 2; j >= 0; --j)
        x[j] -= gam[j + 1] * x[j + 1]; /* :43-46 */
    return 0;
}

/* t-grid, src/Pricer/kwFd1d.cpp:95-103. */
void kwo_t_grid(double tMax, int64_t tDim, double* t)
{
    const double tMin = 0.0;
    const double dt = (tMax - tMin) / (double)(uint64_t)(tDim - 1);
    for (int j = 0; j < tDim; ++j)
        t[j] = tMin + j * dt;
}

/* TEST HOOK (off by default; with 0 the restatement is bit-for-bit the reference): model "the same reference
 * built against another libm".  sinh() and exp() are not correctly rounded in any libm -- glibc's are < 1 ulp,
 * CUDA's <= 2 ulp -- so two faithful implementations of src/Pricer/kwFd1d.cpp:120-139 legitimately differ in the
 * last bits of x_j and of the payoff.  With a jitter of n, every sinh() / exp() result is moved by 0..n ulp in a
 * pseudo-random direction.  tests/ use the resulting price change as the measured libm sensitivity of a grid
 * shape: where it exceeds 1e-9 (few time steps on a very fine grid, dt/dx^2 in the thousands) the reference does
 * not define its own prices to 1e-9 and the GPU-vs-oracle bar is that sensitivity (DESIGN.md "Parity budget"). */
static int g_libm_jitter = 0;
void kwo_set_libm_jitter(int ulps) { g_libm_jitter = ulps; }
static double libm_jitter(double v, uint64_t key)
{
    if (!g_libm_jitter) return v;
    key += 0x9e3779b97f4a7c15ull;
    key = (key ^ (key >> 30)) * 0xbf58476d1ce4e5b9ull;
    key = (key ^ (key >> 27)) * 0x94d049bb133111ebull;
    key ^= key >> 31;
    const int steps = (int)(key % (uint64_t)(g_libm_jitter + 1));
    const double to = (key >> 40) & 1 ? INFINITY : -INFINITY;
    for (int i = 0; i < steps; ++i) v = nextafter(v, to);
    return v;
}

/* x-grid, src/Pricer/kwFd1d.cpp:105-125: log-moneyness sinh grid centred on
 * the strike, independent of s and k. */
void kwo_x_grid(double z, double t, double density, double scale, int64_t xDim, double* x)
{
    const double xMid = 0.;
    const double xMin = xMid - scale * z * sqrt(t);
    const double xMax = xMid + scale * z * sqrt(t);
    const double yMin = asinh((xMin - xMid) / density);
    const double yMax = asinh((xMax - xMid) / density);
    const double dy = 1. / (double)(uint64_t)(xDim - 1);
    for (int j = 0; j < xDim; ++j) {
        const double yj = j * dy;
        x[j] = xMid + density * libm_jitter(sinh(yMin * (1.0 - yj) + yMax * yj), (uint64_t)j);
    }
}

/* payoff in units of the strike, src/Pricer/kwFd1d.cpp:127-139. */
void kwo_payoff(int w, int64_t xDim, const double* x, double* v)
{
    for (int j = 0; j < xDim; ++j) {
        if (w < 0) {
            const double p = 1. - libm_jitter(exp(x[j]), (uint64_t)j + 0x100000u);
            v[j] = 0 < p ? p : 0; /* std::max<f64>(0, p): returns 0 unless 0 < p */
        } else {
            const double p = libm_jitter(exp(x[j]), (uint64_t)j + 0x100000u) - 1.;
            v[j] = 0 < p ? p : 0;
        }
    }
}

/* Per-thread work arrays: the reference keeps (7*xDim + 4*tDim) doubles per
 * PDE (src/Math/kwFd1d.cpp:19-27); the arithmetic only ever touches one PDE's
 * rows at a time, so one set per worker thread gives identical results. */
typedef struct {
    double *t, *x, *v0, *bl, *b, *bu, *w, *v, *gam;
} kwo_work;

static int work_alloc(kwo_work* k, int64_t tDim, int64_t xDim)
{
    k->t = (double*)malloc(sizeof(double) * (size_t)tDim);
    double* p = (double*)malloc(sizeof(double) * (size_t)xDim * 8);
    if (!k->t || !p)
        return -1;
    k->x = p;
    k->v0 = p + xDim;
    k->bl = p + 2 * xDim;
    k->b = p + 3 * xDim;
    k->bu = p + 4 * xDim;
    k->w = p + 5 * xDim;
    k->v = p + 6 * xDim;
    k->gam = p + 7 * xDim;
    return 0;
}

static void work_free(kwo_work* k)
{
    free(k->t);
    free(k->x);
}

/* Fd1d::solveOne, src/Math/kwFd1d.cpp:61-136.  theta = 0.5 (src/Math/kwFd1d.h:31).
 * a0/ax/axx are constant along t (src/Math/kwFd1d.cpp:32-36).  On return
 * v holds the t=0 solution on the x grid. */
void kwo_solve_one(int64_t tDim, int64_t xDim, double a0, double ax, double axx, int earlyExercise,
                   const double* t, const double* x, const double* v0, double* bl, double* b,
                   double* bu, double* w, double* v, double* gam)
{
    const double theta = 0.5;
    for (int64_t j = 0; j < xDim; ++j)
        v[j] = v0[j]; /* src/Math/kwFd1d.cpp:40-41 */

    for (int ti = (int)tDim - 2; ti >= 0; --ti) {
        const double dt = t[ti + 1] - t[ti]; /* :72 */
        {                                    /* row 0, :73-84 */
            const int xi = 0;
            const double inv_dx = 1. / (x[xi + 1] - x[xi]);
            bl[xi] = 0;
            b[xi] = 1 - theta * dt * (a0 - inv_dx * ax);
            bu[xi] = -theta * dt * (inv_dx * ax);
            w[xi] = (1 + (1 - theta) * dt * a0) * v[xi] +
                    (1 - theta) * dt * (ax * inv_dx) * (v[xi + 1] - v[xi]);
        }
        for (int xi = 1; xi < xDim - 1; ++xi) { /* interior, :86-102 */
            const double inv_dxu = 1. / (x[xi + 1] - x[xi]);
            const double inv_dxm = 1. / (x[xi + 1] - x[xi - 1]);
            const double inv_dxd = 1. / (x[xi] - x[xi - 1]);

            const double inv_dx2u = 2. * inv_dxu * inv_dxm;
            const double inv_dx2m = 2. * inv_dxd * inv_dxu;
            const double inv_dx2l = 2. * inv_dxd * inv_dxm;

            bl[xi] = -theta * dt * (-inv_dxm * ax + inv_dx2l * axx);
            b[xi] = 1 - theta * dt * (a0 - inv_dx2m * axx);
            bu[xi] = -theta * dt * (inv_dxm * ax + inv_dx2u * axx);

            w[xi] = (1 + (1 - theta) * dt * a0) * v[xi] +
                    (1 - theta) * dt * (ax * inv_dxm) * (v[xi + 1] - v[xi - 1]) +
                    (1 - theta) * dt * (axx) *
                        (inv_dx2u * v[xi + 1] - inv_dx2m * v[xi] + inv_dx2l * v[xi - 1]);
        }
        { /* row xDim-1, :104-114 */
            const int xi = (int)xDim - 1;
            const double inv_dx = 1. / (x[xi] - x[xi - 1]);
            bl[xi] = -theta * dt * (-inv_dx * ax);
            b[xi] = 1 - theta * dt * (a0 + inv_dx * ax);
            bu[xi] = 0;
            w[xi] = (1 + (1 - theta) * dt * a0) * v[xi] +
                    (1 - theta) * dt * (ax * inv_dx) * (v[xi] - v[xi - 1]);
        }

        /* :118-125 -- the reference discards the solver's error code. */
        (void)kwo_solve_tridiagonal((int)xDim, bl, b, bu, w, v, gam);

        if (earlyExercise) { /* :129-132 -- last node is NOT projected */
            for (int xi = 0; xi < xDim - 1; ++xi)
                v[xi] = v[xi] < v0[xi] ? v0[xi] : v[xi]; /* std::max(v, v0) */
        }
    }
}

/* Fd1d::value, src/Math/kwFd1d.cpp:139-158.  Returns 0 on success, 1 if x_ is
 * not strictly inside (x[0], x[xDim-1]] (reference error, :151-153). */
int kwo_value(int64_t xDim, const double* x, const double* v, double x_, double* out)
{
    int64_t xi = 0;
    while ((xi < xDim) && (x[xi] < x_))
        ++xi;
    if ((xi == 0) || (xi == xDim))
        return 1;
    *out = ((x[xi] - x_) * v[xi - 1] + (x_ - x[xi - 1]) * v[xi]) / (x[xi] - x[xi - 1]);
    return 0;
}

/* Lexicographic (t,r,q,z,e,w) comparison: src/Pricer/kwFd1d.cpp:33-35. */
static const kwo_option* g_sort_assets; /* qsort has no context arg; guarded by single caller */
static int key_less(const kwo_option* l, const kwo_option* r)
{
    if (l->t != r->t) return l->t < r->t;
    if (l->r != r->r) return l->r < r->r;
    if (l->q != r->q) return l->q < r->q;
    if (l->z != r->z) return l->z < r->z;
    if (l->e != r->e) return l->e < r->e;
    return l->w < r->w;
}
static int idx_cmp(const void* a, const void* b)
{
    const uint32_t l = *(const uint32_t*)a, r = *(const uint32_t*)b;
    if (key_less(&g_sort_assets[l], &g_sort_assets[r])) return -1;
    if (key_less(&g_sort_assets[r], &g_sort_assets[l])) return 1;
    return l < r ? -1 : (l > r); /* deterministic tie-break; the PDE depends on the key only */
}

int kwo_max_threads(void);

/* One PDE per job, as the reference's thread pool (src/Math/kwFd1d.cpp:45-55);
 * here plain pthreads pulling PDE indices from a shared counter. */
typedef struct {
    const kwo_option* assets;
    double* prices;
    const uint32_t *pde2asset, *start, *members;
    long long m;
    double density, scale;
    int64_t tDim, xDim;
    long long next;
    int failed;
    long long fail_option;
    pthread_mutex_t mu;
} kwo_job;

static void* worker(void* arg)
{
    kwo_job* J = (kwo_job*)arg;
    kwo_work k;
    if (work_alloc(&k, J->tDim, J->xDim)) {
        pthread_mutex_lock(&J->mu);
        if (!J->failed) J->failed = 3;
        pthread_mutex_unlock(&J->mu);
        return NULL;
    }
    for (;;) {
        const long long p = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (p >= J->m) break;
        const kwo_option* a = &J->assets[J->pde2asset[p]];
        /* PDE coefficients, src/Pricer/kwFd1d.cpp:68-86 */
        const double a0 = -a->r;
        const double ax = a->r - a->q - a->z * a->z / 2;
        const double axx = a->z * a->z / 2;

        kwo_t_grid(a->t, J->tDim, k.t);
        kwo_x_grid(a->z, a->t, J->density, J->scale, J->xDim, k.x);
        kwo_payoff(a->w, J->xDim, k.x, k.v0);
        kwo_solve_one(J->tDim, J->xDim, a0, ax, axx, a->e != 0, k.t, k.x, k.v0, k.bl, k.b, k.bu,
                      k.w, k.v, k.gam);

        /* fill prices, src/Pricer/kwFd1d.cpp:146-157 */
        for (uint32_t q = J->start[p]; q < J->start[p + 1]; ++q) {
            const uint32_t i = J->members[q];
            double price_ = 0;
            const double xq = log(J->assets[i].s / J->assets[i].k);
            if (kwo_value(J->xDim, k.x, k.v, xq, &price_)) {
                pthread_mutex_lock(&J->mu);
                if (J->failed != 1 || (long long)i < J->fail_option) {
                    J->failed = 1;
                    J->fail_option = i;
                }
                pthread_mutex_unlock(&J->mu);
                J->prices[i] = NAN;
            } else {
                J->prices[i] = J->assets[i].k * price_;
            }
        }
    }
    work_free(&k);
    return NULL;
}

/* Fd1d_Pricer::price, src/Pricer/kwFd1d.cpp:21-160 (mode "FD1D").
 * compress != 0: one PDE per (t,r,q,z,e,w) chain as the reference (:28-65);
 * compress == 0: one PDE per option (same prices, because a PDE's solution
 * depends on the chain key only -- used to bound memory/time in tests).
 * nthreads <= 0: all cores.  Returns 0 on success; 1 + message on the
 * reference's interpolation error (:154-155); n == 0 leaves prices untouched
 * (:24-26). */
int kwo_fd1d_price(const kwo_option* assets, size_t n, double density, double scale, int64_t tDim,
                   int64_t xDim, int compress, int nthreads, double* prices, uint32_t* n_pde_out,
                   char* err, size_t errlen)
{
    if (n_pde_out) *n_pde_out = 0;
    if (n == 0)
        return 0;
    if (tDim < 2 || xDim < 3) {
        set_err(err, errlen, "kwo_fd1d_price: grid too small");
        return 2;
    }

    uint32_t* asset2pde = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t* pde2asset = (uint32_t*)malloc(sizeof(uint32_t) * n);
    size_t m = 0;
    if (compress) {
        uint32_t* sorted = (uint32_t*)malloc(sizeof(uint32_t) * n);
        for (size_t i = 0; i < n; ++i) sorted[i] = (uint32_t)i;
        g_sort_assets = assets;
        qsort(sorted, n, sizeof(uint32_t), idx_cmp);
        asset2pde[sorted[0]] = 0;
        pde2asset[m++] = sorted[0];
        for (size_t i = 1, k = 0; i < n; ++i) {
            const uint32_t l = sorted[i - 1], r = sorted[i];
            const int equal = key_less(&assets[l], &assets[r]) == key_less(&assets[r], &assets[l]);
            if (!equal) {
                k += 1;
                pde2asset[m++] = r;
            }
            asset2pde[r] = (uint32_t)k;
        }
        free(sorted);
    } else {
        for (size_t i = 0; i < n; ++i) {
            asset2pde[i] = (uint32_t)i;
            pde2asset[i] = (uint32_t)i;
        }
        m = n;
    }
    if (n_pde_out) *n_pde_out = (uint32_t)m;

    /* CSR pde -> options so each worker interpolates its own chain */
    uint32_t* start = (uint32_t*)calloc(m + 1, sizeof(uint32_t));
    uint32_t* members = (uint32_t*)malloc(sizeof(uint32_t) * n);
    for (size_t i = 0; i < n; ++i) start[asset2pde[i] + 1]++;
    for (size_t p = 0; p < m; ++p) start[p + 1] += start[p];
    {
        uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * m);
        memcpy(fill, start, sizeof(uint32_t) * m);
        for (size_t i = 0; i < n; ++i) members[fill[asset2pde[i]]++] = (uint32_t)i;
        free(fill);
    }

    kwo_job job;
    job.assets = assets; job.prices = prices; job.pde2asset = pde2asset; job.start = start;
    job.members = members; job.m = (long long)m; job.density = density; job.scale = scale;
    job.tDim = tDim; job.xDim = xDim; job.next = 0; job.failed = 0; job.fail_option = -1;
    pthread_mutex_init(&job.mu, NULL);
    if (nthreads <= 0) nthreads = kwo_max_threads();
    if ((long long)nthreads > (long long)m) nthreads = (int)m;
    if (nthreads <= 1) {
        worker(&job);
    } else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
        int started = 0;
        for (int i = 0; i < nthreads; ++i)
            if (pthread_create(&th[started], NULL, worker, &job) == 0) started++;
        if (started == 0) worker(&job);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&job.mu);
    const int failed = job.failed;
    const long long fail_option = job.fail_option;

    free(asset2pde);
    free(pde2asset);
    free(start);
    free(members);

    if (failed == 1) {
        char msg[160];
        snprintf(msg, sizeof msg, "Fd1d_Pricer::price Fd1d::value: x not in range (option %lld)",
                 fail_option);
        set_err(err, errlen, msg);
        return 1;
    }
    if (failed) {
        set_err(err, errlen, "kwo_fd1d_price: out of memory");
        return 3;
    }
    return 0;
}

/* Full solution vector of ONE option's PDE on its x grid (for kernel-level
 * parity tests): x_out[xDim], v_out[xDim]. */
int kwo_fd1d_solution(const kwo_option* a, double density, double scale, int64_t tDim, int64_t xDim,
                      double* x_out, double* v_out)
{
    kwo_work k;
    if (work_alloc(&k, tDim, xDim)) return 3;
    const double a0 = -a->r;
    const double ax = a->r - a->q - a->z * a->z / 2;
    const double axx = a->z * a->z / 2;
    kwo_t_grid(a->t, tDim, k.t);
    kwo_x_grid(a->z, a->t, density, scale, xDim, k.x);
    kwo_payoff(a->w, xDim, k.x, k.v0);
    kwo_solve_one(tDim, xDim, a0, ax, axx, a->e != 0, k.t, k.x, k.v0, k.bl, k.b, k.bu, k.w, k.v,
                  k.gam);
    memcpy(x_out, k.x, sizeof(double) * (size_t)xDim);
    memcpy(v_out, k.v, sizeof(double) * (size_t)xDim);
    work_free(&k);
    return 0;
}

/* kw::cdfNormal, src/Math/kwMath.cpp:9-13. */
static double cdf_normal(double x)
{
    return 0.5 * (1 + erf(x / 1.4142135623730950488016887242097 /* std::numbers::sqrt2 */));
}

/* BlackScholes_Pricer::priceOne, src/Pricer/kwBlackScholes.cpp:27-50:
 * NaN for American options (:30-33). */
void kwo_bs_price(const kwo_option* assets, size_t n, double* prices)
{
    for (size_t i = 0; i < n; ++i) {
        const kwo_option* a = &assets[i];
        if (a->e) {
            prices[i] = NAN;
            continue;
        }
        const double k = a->k, q = a->q, r = a->r, s = a->s, t = a->t, z = a->z;
        const int w = a->w;
        const double zt = z * sqrt(t);
        const double d1 = 1 / zt * (log(s / k) + (r - q + 0.5 * z * z) * t);
        const double d2 = d1 - zt;
        prices[i] = w * (s * cdf_normal(w * d1) * exp(-q * t) - k * cdf_normal(w * d2) * exp(-r * t));
    }
}

/* Fd1d_BlackScholes_Pricer::price, src/Pricer/kwFd1d_BlackScholes.cpp:15-43:
 * FD(as given) + (BS_euro - FD_euro). */
int kwo_fd1d_bs_price(const kwo_option* assets, size_t n, double density, double scale, int64_t tDim,
                      int64_t xDim, int compress, int nthreads, double* prices, char* err,
                      size_t errlen)
{
    if (n == 0) return 0;
    int rc = kwo_fd1d_price(assets, n, density, scale, tDim, xDim, compress, nthreads, prices, NULL,
                            err, errlen);
    if (rc) return rc;
    kwo_option* euro = (kwo_option*)malloc(sizeof(kwo_option) * n);
    double* fd = (double*)malloc(sizeof(double) * n);
    double* bs = (double*)malloc(sizeof(double) * n);
    memcpy(euro, assets, sizeof(kwo_option) * n);
    for (size_t i = 0; i < n; ++i) euro[i].e = 0;
    rc = kwo_fd1d_price(euro, n, density, scale, tDim, xDim, compress, nthreads, fd, NULL, err, errlen);
    if (!rc) {
        kwo_bs_price(euro, n, bs);
        for (size_t i = 0; i < n; ++i) prices[i] += bs[i] - fd[i];
    }
    free(euro);
    free(fd);
    free(bs);
    return rc;
}

int kwo_max_threads(void)
{
    const long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

size_t kwo_sizeof_option(void) { return sizeof(kwo_option); }
