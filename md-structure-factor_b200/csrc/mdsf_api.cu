// libmdsf.so -- host side of the C ABI declared in include/mdsf.h.
// One handle = one GPU, three streams (H2D copy, compute, D2H write-back) and a double-buffered
// coordinate staging area, so the copy of batch b+1 overlaps the kernels of batch b.
#include <cub/cub.cuh>
#include <cuda.h>               // green-context TYPES only: the driver entry points are fetched at run time
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mdsf.h"
#include "mdsf_common.cuh"
#include "mdsf_fft.cuh"
#include "mdsf_prep.cuh"
#include "mdsf_splat.cuh"
#include "mdsf_scatter.cuh"

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(MDSF_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CF(call)                                                                              \
    do {                                                                                      \
        cufftResult r_ = (call);                                                              \
        if (r_ != CUFFT_SUCCESS) return fail(MDSF_ECUDA, "%s failed: cufft error %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); \
    } while (0)

#ifndef MDSF_FFT3_MINB_Y
#define MDSF_FFT3_MINB_Y 2     // CTAs per SM of the three-stage y pass (measured: 2 -> 4.87 ms, 3 -> 4.95 ms on c3)
#endif
#ifndef MDSF_FFT3_MINB_X
#define MDSF_FFT3_MINB_X 2     // ... of the x pass (register accumulators: 116 registers)
#endif
static const int kSlots = 2;
static const int kMaxSmem = 227 * 1024;

struct AxisPlan {
    FftPlan plan{};
    bool native = false;
    double2* d_tw = nullptr;
    double2* d_tw16 = nullptr;  // n = 256: inter-stage twiddles of the 16x16 split in [k][n2] order (w^(n2 k) at k*16 + n2)
    int* d_rev = nullptr;     // frequency index -> position
};

struct mdsf_handle {
    mdsf_config cfg{};
    GridParams gp{};
    TypeTable tt{};
    int device = 0;
    int nsm = 148;
    size_t csize = 4;                 // sizeof coordinate dtype
    bool native_fft = false;
    int F = 2;                        // frames per batch
    long long ncell = 0;
    // streams / events
    cudaStream_t s_copy = nullptr, s_comp = nullptr, s_back = nullptr;
    cudaEvent_t ev_h2d[kSlots]{}, ev_free[kSlots]{}, ev_prep[kSlots]{}, ev_back[kSlots]{};
    bool slot_used[kSlots]{};
    int next_slot = 0;
    // device buffers
    double *d_amp = nullptr, *d_two = nullptr, *d_ctab = nullptr, *d_tables = nullptr;     // d_tables: alias of the current set
    int *d_halfw = nullptr, *d_ctab_off = nullptr;
    unsigned* d_toff = nullptr;
    std::vector<double> two_host;
    int logS = 4, zstage = 1023;
    bool ez_global = false;
    // scatter (fixed-point, slab-pipelined) splat mode
    bool scatter = false, tile_atomic = false;
    static const int kMarks = 16;             // mdsf_input_mark / mdsf_input_wait tickets
    cudaEvent_t mark_copy[kMarks] = {}, mark_back[kMarks] = {};
    int64_t next_mark = 0;
    int want_mode = 0;                // 0 auto, 1 owner, 2 scatter, 3 tile (shared-memory fixed-point atomics)
    SlabParams sp{};
    unsigned long long* d_acc = nullptr;
    unsigned *d_slab_count = nullptr, *d_slab_start = nullptr, *d_slab_cursor = nullptr, *d_entries = nullptr;
    unsigned *d_step_start = nullptr, *d_ctl = nullptr;
    int ring = 3;
    int zfast = 0;                    // 16 / 8: Nz = R*R handled by the two-stage z pass with natural-order output
    long long entries_cap = 0;
    int zcol = 16;
    size_t acc_cells = 0, l2_window = 0;
    float l2_ratio = 1.0f;
    int pipe_grid = 0;
    int* d_type = nullptr;
    void* d_stage[kSlots]{};
    // outputs of the prep/bin stage; two sets so that prep+bin of batch b+1 (stream s_prep) overlap the
    // splat/FFT kernels of batch b (stream s_comp).  The d_* names below alias the set of the current batch.
    struct PrepSet {
        AtomRec* recs = nullptr;
        unsigned *cnt = nullptr, *off = nullptr, *keys[2]{}, *vals[2]{}, *tile_start = nullptr;
        unsigned* counter = nullptr;      // direct binning: [nkeys+1] list lengths, then [nkeys+1] cursors
        uint4* prec = nullptr;            // direct binning: 16-byte pair records in list order (splat_zfft_kernel, PREC)
        double* tables = nullptr;
        void* cub = nullptr;
        cudaEvent_t ev_binned = nullptr, ev_consumed = nullptr;
        bool used = false;
    } sets[2];
    int nsets = 1;
    long long batch_counter = 0;
    cudaStream_t s_prep = nullptr;
    AtomRec* d_recs = nullptr;
    unsigned *d_cnt = nullptr, *d_off = nullptr;
    unsigned *d_keys[2]{}, *d_vals[2]{};
    unsigned* d_tile_start = nullptr;
    unsigned* d_counter = nullptr;
    void* d_cub = nullptr;
    size_t cub_bytes = 0;
    double2* d_vol = nullptr;
    double2* d_dump = nullptr;
    double* d_P = nullptr;
    double* d_sf = nullptr;
    int* d_err = nullptr;
    int* h_err = nullptr;
    AxisPlan ax[3];
    cufftHandle cufft_plan = 0;
    int cufft_batch = 0;
    // splat launch geometry
    long long natoms = 0;
    long long maxpairs_frame = 0;
    int chunk = 128;
    size_t splat_smem = 0;
    int mono = 0;                     // K1 applies the monoclinic transform of main_gromacs.py:204-207 first
    double mono_sin = 1.0, mono_cos = 0.0;
    int pf_dist = 0;                  // MDSF_PF_DIST = n: splat CTAs warm L2 for the CTA n tiles later; measured slower (c2 5.54 -> 5.77 ms), off
    bool pair_records = true;         // direct binning writes 16-byte pair records (MDSF_PAIR_RECORDS=0: 4-byte payloads + atom records)
    bool direct_bin = false;          // tile mode: counting-sort binning with atomics instead of the stable radix sort
    int tw16_off = 0;                 // byte offset of the cp.async-prefetched stage-1 twiddle table in the splat's shared memory (0 = none)
    int sort_bits = 1;
    // y/x pass geometry
    int Wy = 16, Wx = 16, thr_y = 256, thr_x = 256;
    // bookkeeping
    long long launches = 0, frames_done = 0;
    int last_batch_frames = 0;
    bool timing = false;
    std::vector<cudaEvent_t> tev;     // 6 events per timed batch: h2d start, compute start, binned, splat done, y done, end
    long long timed_batches = 0;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    std::vector<int> halfw_host;
    // overlap mode (MDSF_SM_SPLIT): the issue-bound prep/bin/splat kernels of batch b+1 run on stream s_splat
    // while the HBM-bound y/x passes of batch b run on s_comp, each on its own pair-volume set.  With
    // MDSF_SM_SPLIT = n > 0 the two streams belong to two green contexts that own disjoint SM partitions
    // (n SMs for the splat side, the rest for the passes); -1 overlaps on plain streams.
    int overlap = 0;
    int x_async = 1;                  // MDSF_X_ASYNC: cp.async-prefetched two-stage x pass (default on)
    int y_async = 0;                  // MDSF_Y_ASYNC: same for the y pass
    int part_sms[2] = {0, 0};
    cudaStream_t s_splat = nullptr;
    double2* d_volset[2] = {nullptr, nullptr};
    cudaEvent_t ev_splat[2]{}, ev_volfree[2]{};
    bool vol_used[2]{};
    CUgreenCtx gctx[2] = {nullptr, nullptr};
};

// ------------------------------------------------------------------------------------------
// SM partitions through CUDA green contexts.  libmdsf.so does not link libcuda (it must load on a box
// without a driver, where mdsf_create then fails with MDSF_ECUDA): the few driver entry points are
// resolved through the runtime.
struct GreenApi {
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool ok = false;
};
static GreenApi& green_api() {
    static GreenApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    bool ok = true;
    auto get = [&](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult st = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*fn) ok = false;
    };
    get("cuDeviceGet", (void**)&api.DeviceGet);
    get("cuDeviceGetDevResource", (void**)&api.DeviceGetDevResource);
    get("cuDevSmResourceSplitByCount", (void**)&api.DevSmResourceSplitByCount);
    get("cuDevResourceGenerateDesc", (void**)&api.DevResourceGenerateDesc);
    get("cuGreenCtxCreate", (void**)&api.GreenCtxCreate);
    get("cuGreenCtxDestroy", (void**)&api.GreenCtxDestroy);
    get("cuGreenCtxStreamCreate", (void**)&api.GreenCtxStreamCreate);
    (void)cudaGetLastError();
    api.ok = ok;
    return api;
}
#define DRV(call)                                                                             \
    do {                                                                                      \
        CUresult r_ = (call);                                                                 \
        if (r_ != CUDA_SUCCESS) return fail(MDSF_ECUDA, "%s failed: driver error %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); \
    } while (0)

// two green contexts: partition 0 with (at least) want_a SMs for prep/bin/splat, partition 1 with the
// remaining SMs for the y/x passes; one non-blocking stream in each
static int make_partitions(mdsf_handle* h, int want_a) {
    GreenApi& api = green_api();
    if (!api.ok) return fail(MDSF_ECUDA, "MDSF_SM_SPLIT: this driver does not export the green-context entry points");
    if (want_a < 8 || want_a > h->nsm - 8) return fail(MDSF_EINVAL, "MDSF_SM_SPLIT=%d: need 8 <= n <= %d", want_a, h->nsm - 8);
    CU(cudaFree(0));                                     // primary context is active
    CUdevice dev;
    DRV(api.DeviceGet(&dev, h->device));
    CUdevResource all, part[2];
    DRV(api.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned int nb = 1;
    DRV(api.DevSmResourceSplitByCount(&part[0], &nb, &all, &part[1], 0, (unsigned)want_a));
    if (nb != 1 || part[0].type != CU_DEV_RESOURCE_TYPE_SM || part[1].type != CU_DEV_RESOURCE_TYPE_SM || part[1].sm.smCount == 0)
        return fail(MDSF_ECUDA, "MDSF_SM_SPLIT=%d: the driver could not split %u SMs that way", want_a, all.sm.smCount);
    for (int i = 0; i < 2; ++i) {
        CUdevResourceDesc desc;
        DRV(api.DevResourceGenerateDesc(&desc, &part[i], 1));
        DRV(api.GreenCtxCreate(&h->gctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
        h->part_sms[i] = (int)part[i].sm.smCount;
        CUstream st;
        DRV(api.GreenCtxStreamCreate(&st, h->gctx[i], CU_STREAM_NON_BLOCKING, 0));
        (i == 0 ? h->s_splat : h->s_comp) = (cudaStream_t)st;
    }
    return MDSF_OK;
}

// ------------------------------------------------------------------------------------------
static bool factorize(int n, FftPlan& plan, int max_log2 = 4) {
    plan.n = n;
    plan.nstages = 0;
    int e = 0;
    while (n % 2 == 0) { n /= 2; ++e; }
    if (e > 0) {
        const int count = (e + max_log2 - 1) / max_log2, base = e / count, rem = e % count;
        for (int i = 0; i < count; ++i) plan.radix[plan.nstages++] = 1 << (base + (i >= count - rem ? 1 : 0));
    }
    const int odd[] = {3, 5, 7, 11, 13};
    for (int p : odd)
        while (n % p == 0) {
            if (plan.nstages >= MDSF_MAX_RADIX_STAGES) return false;
            plan.radix[plan.nstages++] = p;
            n /= p;
        }
    return n == 1 && plan.nstages > 0;
}

static int digit_position(int k, int n, const FftPlan& plan, int stage) {
    if (stage >= plan.nstages) return 0;
    const int r = plan.radix[stage], m = n / r;
    return (k % r) * m + digit_position(k / r, m, plan, stage + 1);
}

static int build_axis(AxisPlan& ax, int n, bool want_native, int max_log2) {
    std::vector<int> rev(n);
    ax.native = want_native && factorize(n, ax.plan, max_log2);
    if (ax.native) {
        std::vector<double2> tw(n);
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int j = 0; j < n; ++j) {
            const long double a = two_pi * (long double)j / (long double)n;
            tw[j].x = (double)cosl(a);
            tw[j].y = (double)(-sinl(a));
        }
        for (int k = 0; k < n; ++k) rev[k] = digit_position(k, n, ax.plan, 0);
        CU(cudaMalloc(&ax.d_tw, sizeof(double2) * n));
        CU(cudaMemcpy(ax.d_tw, tw.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
        if (n == 256) {
            std::vector<double2> t16(256);
            for (int i = 0; i < 256; ++i) t16[i] = tw[(i >> 4) * (i & 15)];
            CU(cudaMalloc(&ax.d_tw16, sizeof(double2) * 256));
            CU(cudaMemcpy(ax.d_tw16, t16.data(), sizeof(double2) * 256, cudaMemcpyHostToDevice));
        }
    } else {
        ax.plan.n = n;
        ax.plan.nstages = 0;
        for (int k = 0; k < n; ++k) rev[k] = k;
    }
    CU(cudaMalloc(&ax.d_rev, sizeof(int) * n));
    CU(cudaMemcpy(ax.d_rev, rev.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    return MDSF_OK;
}

static int grid_for(long long n, int threads, int nsm) {
    long long b = (n + threads - 1) / threads;
    long long cap = (long long)nsm * 16;
    return (int)std::max(1LL, std::min(b, cap));
}

// ------------------------------------------------------------------------------------------
extern "C" int mdsf_abi_version(void) { return MDSF_ABI_VERSION; }
extern "C" const char* mdsf_last_error(void) { return g_err.c_str(); }

extern "C" int mdsf_create(const mdsf_config* cfg, mdsf_handle** out) {
    if (!cfg || !out) return fail(MDSF_EINVAL, "null argument");
    if (cfg->abi_version != MDSF_ABI_VERSION) return fail(MDSF_EINVAL, "ABI version mismatch (%d != %d)", cfg->abi_version, MDSF_ABI_VERSION);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(MDSF_ECUDA, "no CUDA device available (%s); libmdsf has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(MDSF_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);
    for (int d = 0; d < 3; ++d) {
        if (cfg->n[d] < 2 || cfg->n[d] % 2) return fail(MDSF_EINVAL, "grid size n[%d]=%d must be even and >= 2", d, cfg->n[d]);
        if (cfg->nborder > cfg->n[d]) return fail(MDSF_EINVAL, "Nborder %d exceeds grid size %d: the reference's fold slices are ill-formed here", cfg->nborder, cfg->n[d]);
        if (!(cfg->dr[d] > 0)) return fail(MDSF_EINVAL, "dr[%d] must be positive", d);
    }
    if (cfg->ntypes < 1 || !cfg->amp || !cfg->two_sig2 || !cfg->halfw) return fail(MDSF_EINVAL, "type tables missing");
    if (cfg->nborder < 0) return fail(MDSF_EINVAL, "negative Nborder");
    if (cfg->ntypes > 1024) return fail(MDSF_EINVAL, "more than 1024 distinct labels");
    for (int t = 0; t < cfg->ntypes * 3; ++t)
        if (2 * cfg->halfw[t] > MDSF_MAX_STAMP) return fail(MDSF_EINVAL, "stamp of %d cells exceeds %d", 2 * cfg->halfw[t], MDSF_MAX_STAMP);
    for (int t = 0; t < cfg->ntypes * 3; ++t)
        if (cfg->halfw[t] < 0 || cfg->halfw[t] > cfg->nborder) return fail(MDSF_EINVAL, "half width %d outside [0, Nborder=%d]", cfg->halfw[t], cfg->nborder);
    if (cfg->coord_dtype != MDSF_F32 && cfg->coord_dtype != MDSF_F64) return fail(MDSF_EINVAL, "bad coord_dtype");
    if (cfg->arith_dtype != MDSF_F32 && cfg->arith_dtype != MDSF_F64) return fail(MDSF_EINVAL, "bad arith_dtype");
    if (cfg->coord_dtype == MDSF_F64 && cfg->arith_dtype == MDSF_F32) return fail(MDSF_EINVAL, "float64 coordinates never promote to float32");

    CU(cudaSetDevice(cfg->device));
    mdsf_handle* h = new mdsf_handle();
    h->cfg = *cfg;
    h->device = cfg->device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    h->nsm = prop.multiProcessorCount;
    h->csize = cfg->coord_dtype == MDSF_F32 ? 4 : 8;
    GridParams& gp = h->gp;
    for (int d = 0; d < 3; ++d) { gp.n[d] = cfg->n[d]; gp.dr[d] = cfg->dr[d]; gp.box[d] = cfg->box[d]; }
    for (int i = 0; i < 9; ++i) gp.u[i] = cfg->ucell[i];
    gp.nb = cfg->nborder;
    h->want_mode = cfg->splat_mode;
    if (getenv("MDSF_SPLAT_MODE")) h->want_mode = atoi(getenv("MDSF_SPLAT_MODE"));
    if (h->want_mode < 0 || h->want_mode > 3) return fail(MDSF_EINVAL, "splat mode must be 0 (auto), 1 (owner), 2 (scatter) or 3 (tile)");
    h->tile_atomic = h->want_mode == 3 || (h->want_mode == 0 && !getenv("MDSF_AUTO_OWNER"));
    gp.debug_skip = getenv("MDSF_SPLAT_SKIP") ? atoi(getenv("MDSF_SPLAT_SKIP")) : 0;
    gp.fold_mode = cfg->fold_mode;
    // z decouples when ucell[2][0]=ucell[2][1]=0 (b_z feeds only c_2) and ucell[0][2]=ucell[1][2]=0
    gp.separable = (gp.u[6] == 0.0 && gp.u[7] == 0.0 && gp.u[2] == 0.0 && gp.u[5] == 0.0) ? 1 : 0;
    gp.cxx = gp.u[0] * gp.u[0] + gp.u[1] * gp.u[1];
    gp.cyy = gp.u[3] * gp.u[3] + gp.u[4] * gp.u[4];
    gp.gxy = gp.u[0] * gp.u[3] + gp.u[1] * gp.u[4];
    gp.czz = gp.u[8] * gp.u[8];
    h->ncell = (long long)gp.n[0] * gp.n[1] * gp.n[2];

    // ---- FFT plans
    const bool want_native = cfg->fft_mode != MDSF_FFT_CUFFT;
    for (int d = 0; d < 3; ++d) {
        const char* env = getenv(d == 2 ? "MDSF_RADIX_LOG2_Z" : "MDSF_RADIX_LOG2_XY");
        int max_log2 = env ? atoi(env) : 4;
        if (max_log2 < 1 || max_log2 > 4) max_log2 = 4;
        int rc = build_axis(h->ax[d], gp.n[d], want_native, max_log2);
        if (rc) return rc;
    }
    h->native_fft = h->ax[0].native && h->ax[1].native && h->ax[2].native;
    if (gp.n[2] > 4096 || gp.n[1] > 2048 || gp.n[0] > 2048) h->native_fft = false;
    if (!h->native_fft) {
        if (cfg->fft_mode == MDSF_FFT_NATIVE)
            return fail(MDSF_EINVAL, "grid %dx%dx%d has a prime factor > 13 (or an axis too long); native FFT unavailable", gp.n[0], gp.n[1], gp.n[2]);
        for (int d = 0; d < 3; ++d) {   // library path works in natural order on every axis
            std::vector<int> ident(gp.n[d]);
            for (int k = 0; k < gp.n[d]; ++k) ident[k] = k;
            CU(cudaMemcpy(h->ax[d].d_rev, ident.data(), sizeof(int) * gp.n[d], cudaMemcpyHostToDevice));
            h->ax[d].native = false;
        }
    }
    // z-column padding in shared memory: keep the last stage's stride odd
    gp.pad_shift = 31;
    if (h->native_fft) {
        const int last = h->ax[2].plan.radix[h->ax[2].plan.nstages - 1];
        if (last == 16) gp.pad_shift = 4; else if (last == 8) gp.pad_shift = 3; else if (last == 4) gp.pad_shift = 2;
    }
    gp.nzp = (gp.n[2] + (gp.pad_shift < 31 ? (gp.n[2] >> gp.pad_shift) : 0)) | 1;   // odd: columns start in different banks

    // ---- splat tile
    int ncol;
    if (cfg->tile_x > 0 && cfg->tile_y > 0) {
        gp.tx = cfg->tile_x; gp.ty = cfg->tile_y;
        const int nc = gp.tx * gp.ty;
        if (nc < 4 || nc > MDSF_MAX_TILE_COLS || (nc & (nc - 1))) return fail(MDSF_EINVAL, "tile %dx%d: the column count must be 4, 8, 16 or 32", gp.tx, gp.ty);
    } else {
        ncol = 32;
        while (ncol > 4 && (size_t)2 * ncol * gp.nzp * 8 > 80 * 1024) ncol >>= 1;
        if ((size_t)2 * ncol * gp.nzp * 8 > 200 * 1024) return fail(MDSF_EINVAL, "grid too long in z (%d) for the column-tile splat", gp.n[2]);
        const int txs[6] = {1, 1, 2, 2, 4, 4}, tys[6] = {1, 2, 2, 4, 4, 8};
        int l = 0;
        while ((1 << l) < ncol) ++l;
        gp.tx = txs[l]; gp.ty = tys[l];
    }
    if ((gp.tx & (gp.tx - 1)) || (gp.ty & (gp.ty - 1))) return fail(MDSF_EINVAL, "tile sizes must be powers of two");
    gp.nslab = 128 / (gp.tx * gp.ty);
    gp.zs = (gp.n[2] + gp.nslab - 1) / gp.nslab;
    if (h->tile_atomic) { gp.nslab = 1; gp.zs = gp.n[2]; }     // no owners, no z slabs: one list per tile
    gp.ntx = (gp.n[0] + gp.tx - 1) / gp.tx;
    gp.nty = (gp.n[1] + gp.ty - 1) / gp.ty;

    // ---- pipeline shape
    int want_split = getenv("MDSF_SM_SPLIT") ? atoi(getenv("MDSF_SM_SPLIT")) : 0;
    if (want_split != 0 && h->native_fft && h->want_mode != 2) h->overlap = 1;

    // ---- batch size
    int F = cfg->batch_frames;
    if (F <= 0) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const double per_pair = (double)h->ncell * 16.0 * (cfg->keep_density ? 2 : 1);
        const double budget = std::min(8.0e9 * (h->overlap ? 2 : 1), (double)free_b * 0.25) / (h->overlap ? 2 : 1);
        int pairs = (int)std::max(1.0, std::floor(budget / per_pair));
        F = 2 * std::min(pairs, MDSF_MAX_BATCH / 2);
    }
    if (F < 2) F = 2;
    if (F % 2) ++F;
    if (F > MDSF_MAX_BATCH) F = MDSF_MAX_BATCH;
    h->F = F;

    // ---- type tables
    const int nt = cfg->ntypes;
    CU(cudaMalloc(&h->d_amp, sizeof(double) * nt));
    CU(cudaMalloc(&h->d_two, sizeof(double) * nt));
    CU(cudaMalloc(&h->d_halfw, sizeof(int) * nt * 3));
    CU(cudaMemcpy(h->d_amp, cfg->amp, sizeof(double) * nt, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_two, cfg->two_sig2, sizeof(double) * nt, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_halfw, cfg->halfw, sizeof(int) * nt * 3, cudaMemcpyHostToDevice));
    {
        double amax = 0;
        for (int t = 0; t < nt; ++t) amax = std::max(amax, cfg->amp[t]);
        int e = 0;
        (void)std::frexp(amax, &e);                       // amax < 2^e
        gp.fx_scale = std::ldexp(1.0, 52 - e);
        gp.fx_inv = std::ldexp(1.0, e - 52);
    }
    h->halfw_host.assign(cfg->halfw, cfg->halfw + nt * 3);
    h->two_host.assign(cfg->two_sig2, cfg->two_sig2 + nt);
    h->tt.amp = h->d_amp; h->tt.two_sig2 = h->d_two; h->tt.halfw = h->d_halfw;
    h->tt.ctab = nullptr; h->tt.ctab_off = nullptr; h->tt.toff = nullptr;
    if (gp.separable && gp.gxy != 0.0) {
        // cross-term table of every type: C[i][j] = exp(-2 gxy dx dy i j / (2 sigma^2))
        std::vector<int> off(nt);
        std::vector<double> ctab;
        for (int t = 0; t < nt; ++t) {
            off[t] = (int)ctab.size();
            const int ax2 = 2 * cfg->halfw[t * 3], ay2 = 2 * cfg->halfw[t * 3 + 1];
            for (int i = 0; i < ax2; ++i)
                for (int j = 0; j < ay2; ++j)
                    ctab.push_back(std::exp(-(2.0 * gp.gxy * gp.dr[0] * gp.dr[1] * (double)i * (double)j) / cfg->two_sig2[t]));
        }
        if (ctab.empty()) ctab.push_back(1.0);
        CU(cudaMalloc(&h->d_ctab, sizeof(double) * ctab.size()));
        CU(cudaMalloc(&h->d_ctab_off, sizeof(int) * nt));
        CU(cudaMemcpy(h->d_ctab, ctab.data(), sizeof(double) * ctab.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_ctab_off, off.data(), sizeof(int) * nt, cudaMemcpyHostToDevice));
        h->tt.ctab = h->d_ctab; h->tt.ctab_off = h->d_ctab_off;
    }
    h->cfg.amp = nullptr; h->cfg.two_sig2 = nullptr; h->cfg.halfw = nullptr;   // caller-owned, not kept

    // ---- volumes
    const int npairs = F / 2;
    CU(cudaMalloc(&h->d_vol, sizeof(double2) * h->ncell * npairs));
    h->d_volset[0] = h->d_vol;
    if (h->overlap) CU(cudaMalloc(&h->d_volset[1], sizeof(double2) * h->ncell * npairs));
    if (cfg->keep_density) CU(cudaMalloc(&h->d_dump, sizeof(double2) * h->ncell * npairs));
    CU(cudaMalloc(&h->d_P, sizeof(double) * h->ncell));
    CU(cudaMemset(h->d_P, 0, sizeof(double) * h->ncell));
    CU(cudaMalloc(&h->d_sf, sizeof(double) * (long long)gp.n[0] * gp.n[1] * (gp.n[2] / 2 + 1)));
    CU(cudaMalloc(&h->d_err, sizeof(int)));
    CU(cudaMemset(h->d_err, 0, sizeof(int)));
    CU(cudaMallocHost(&h->h_err, sizeof(int)));
    *h->h_err = 0;

    CU(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    if (h->overlap && want_split > 0) {
        int rc = make_partitions(h, want_split);
        if (rc) return rc;
    } else {
        CU(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
        if (h->overlap) CU(cudaStreamCreateWithFlags(&h->s_splat, cudaStreamNonBlocking));
    }
    if (h->overlap)
        for (int v = 0; v < 2; ++v) {
            CU(cudaEventCreateWithFlags(&h->ev_splat[v], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_volfree[v], cudaEventDisableTiming));
        }
    CU(cudaStreamCreateWithFlags(&h->s_back, cudaStreamNonBlocking));
    {
        // prep+bin of batch b+1 run underneath the HBM-bound y/x passes of batch b; at equal priority their small
        // kernels queue behind the passes' CTAs and finish after them (a gap before the next splat), so the prep
        // stream gets the highest priority (MDSF_PREP_PRIO=0 restores the plain stream)
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const bool prio = !getenv("MDSF_PREP_PRIO") || atoi(getenv("MDSF_PREP_PRIO")) != 0;
        CU(cudaStreamCreateWithPriority(&h->s_prep, cudaStreamNonBlocking, prio ? hi : 0));
    }
    for (int s = 0; s < kSlots; ++s) {
        CU(cudaEventCreateWithFlags(&h->ev_h2d[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_free[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_prep[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_back[s], cudaEventDisableTiming));
    }
    CU(cudaEventCreate(&h->timer0));
    CU(cudaEventCreate(&h->timer1));

    if (!h->native_fft) {
        int dims[3] = {gp.n[0], gp.n[1], gp.n[2]};
        CF(cufftPlanMany(&h->cufft_plan, 3, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2Z, npairs));
        CF(cufftSetStream(h->cufft_plan, h->s_comp));
        h->cufft_batch = npairs;
    } else {
        // y / x pass tiles: [n][W] complex in shared memory
        // short axes: [n][8] tiles, 128 threads, 4+ CTAs per SM.  Long axes keep the 128-byte rows and take 256
        // threads (two CTAs per SM) or 512 (one) instead of narrowing the tile (c3: x pass 6.8 -> 5.5 ms).
        auto pick = [&](int n, int& W, int& thr, int arrays) {
            const size_t one = (size_t)n * 8 * 8 * arrays + (size_t)n * 16;       // [n][8] tile(s) + twiddles
            W = 8;
            thr = MDSF_PASS_THREADS;
            if (one > 100 * 1024) thr = one <= 110 * 1024 ? 256 : 512;
            while (W > 1 && (size_t)n * W * 8 * arrays + (size_t)n * 16 > (size_t)kMaxSmem - 2048) W >>= 1;
        };
        pick(gp.n[1], h->Wy, h->thr_y, 2);
        pick(gp.n[0], h->Wx, h->thr_x, 3);
        if (getenv("MDSF_WY")) h->Wy = atoi(getenv("MDSF_WY"));
        if (getenv("MDSF_WX")) h->Wx = atoi(getenv("MDSF_WX"));
        if (getenv("MDSF_THR_Y")) h->thr_y = atoi(getenv("MDSF_THR_Y"));
        if (getenv("MDSF_THR_X")) h->thr_x = atoi(getenv("MDSF_THR_X"));
        CU(cudaFuncSetAttribute(fft_y_kernel<128, MDSF_PASS_MINBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_y_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_y_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_x_accum_kernel<128, MDSF_PASS_MINBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_x_accum_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_x_accum_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_x_accum_async_kernel<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_x_accum_async_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        h->x_async = getenv("MDSF_X_ASYNC") ? atoi(getenv("MDSF_X_ASYNC")) : 1;     // measured: c2 x pass 1.01 -> 0.92 ms
        CU(cudaFuncSetAttribute(fft3_pass_kernel<16, 16, 3, 3, 256, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft3_pass_kernel<16, 16, 3, 3, 256, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft3_pass_kernel<8, 8, 16, 3, 256, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft3_pass_kernel<8, 8, 16, 3, 256, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_Y, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_X, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_y_async_kernel<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CU(cudaFuncSetAttribute(fft_y_async_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        h->y_async = getenv("MDSF_Y_ASYNC") ? atoi(getenv("MDSF_Y_ASYNC")) : 0;
    }
    CU(cudaFuncSetAttribute(slab_pipeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 20480));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<true, true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<true, true, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<true, true, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<true, true, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<true, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<true, false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<true, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<false, true, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<false, true, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<false, true, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<false, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<false, false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CU(cudaFuncSetAttribute(splat_zfft_kernel<false, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    *out = h;
    return MDSF_OK;
}

extern "C" int mdsf_destroy(mdsf_handle* h) {
    if (!h) return MDSF_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->cufft_plan) cufftDestroy(h->cufft_plan);
    void* bufs[] = {h->d_acc, h->d_slab_count, h->d_slab_start, h->d_slab_cursor, h->d_entries, h->d_step_start, h->d_ctl, h->d_ctab, h->d_ctab_off, h->d_toff, h->d_amp, h->d_two, h->d_halfw, h->d_type, h->d_stage[0], h->d_stage[1],
                    h->d_volset[0], h->d_volset[1], h->d_dump, h->d_P, h->d_sf, h->d_err};
    for (void* b : bufs) if (b) cudaFree(b);
    for (int d = 0; d < 3; ++d) { if (h->ax[d].d_tw) cudaFree(h->ax[d].d_tw); if (h->ax[d].d_tw16) cudaFree(h->ax[d].d_tw16); if (h->ax[d].d_rev) cudaFree(h->ax[d].d_rev); }
    for (auto& ps : h->sets) {
        void* pb[] = {ps.recs, ps.cnt, ps.off, ps.keys[0], ps.keys[1], ps.vals[0], ps.vals[1], ps.tile_start, ps.counter, ps.prec, ps.tables, ps.cub};
        for (void* b : pb) if (b) cudaFree(b);
        if (ps.ev_binned) cudaEventDestroy(ps.ev_binned);
        if (ps.ev_consumed) cudaEventDestroy(ps.ev_consumed);
    }
    if (h->s_prep) cudaStreamDestroy(h->s_prep);
    if (h->h_err) cudaFreeHost(h->h_err);
    for (int s = 0; s < kSlots; ++s) {
        if (h->ev_h2d[s]) cudaEventDestroy(h->ev_h2d[s]);
        if (h->ev_free[s]) cudaEventDestroy(h->ev_free[s]);
        if (h->ev_prep[s]) cudaEventDestroy(h->ev_prep[s]);
        if (h->ev_back[s]) cudaEventDestroy(h->ev_back[s]);
    }
    for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
    for (int i = 0; i < mdsf_handle::kMarks; ++i) {
        if (h->mark_copy[i]) cudaEventDestroy(h->mark_copy[i]);
        if (h->mark_back[i]) cudaEventDestroy(h->mark_back[i]);
    }
    if (h->timer0) cudaEventDestroy(h->timer0);
    if (h->timer1) cudaEventDestroy(h->timer1);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    if (h->s_splat) cudaStreamDestroy(h->s_splat);
    if (h->s_back) cudaStreamDestroy(h->s_back);
    for (int v = 0; v < 2; ++v) {
        if (h->ev_splat[v]) cudaEventDestroy(h->ev_splat[v]);
        if (h->ev_volfree[v]) cudaEventDestroy(h->ev_volfree[v]);
        if (h->gctx[v] && green_api().ok) green_api().GreenCtxDestroy(h->gctx[v]);
    }
    delete h;
    return MDSF_OK;
}

extern "C" int mdsf_set_atoms(mdsf_handle* h, int64_t natoms, const int32_t* type_id) {
    if (!h || !type_id) return fail(MDSF_EINVAL, "null argument");
    if (natoms < 1 || natoms >= MDSF_MAX_ATOMS) return fail(MDSF_EINVAL, "natoms=%lld outside [1, %d)", (long long)natoms, MDSF_MAX_ATOMS);
    if (h->d_type) return fail(MDSF_ESTATE, "atoms already set on this handle");
    CU(cudaSetDevice(h->device));
    const GridParams& g0 = h->gp;
    const int nt = h->cfg.ntypes;
    // worst-case (image, tile) pairs per atom of each type: exact maximum over every admissible cell index
    std::vector<long long> bound(nt);
    int zmax = 2;
    for (int t = 0; t < nt; ++t) {
        long long m[2] = {1, 1};
        for (int d = 0; d < 2; ++d) {
            const int A = h->halfw_host[t * 3 + d], N = g0.n[d], tl = d == 0 ? g0.tx : g0.ty;
            for (int ir = A - g0.nb; ir <= N + g0.nb - A; ++ir) m[d] = std::max<long long>(m[d], stamp_tiles_1d(ir, A, N, tl));
        }
        long long mz = 1;
        {
            const int A = h->halfw_host[t * 3 + 2], N = g0.n[2];
            for (int ir = A - g0.nb; ir <= N + g0.nb - A; ++ir)
                for (int sx = 0; sx <= 1; ++sx)
                    for (int sy = -1; sy <= 1; ++sy) {
                        int a1, a2, a3, a4;
                        const unsigned sm = image_slabmask(ir, A, sx, sy, N, g0.nb, g0.fold_mode, g0.zs, a1, a2, a3, a4);
                        mz = std::max<long long>(mz, __builtin_popcount(sm));
                    }
        }
        bound[t] = m[0] * m[1] * mz;
        zmax = std::max(zmax, 2 * h->halfw_host[t * 3 + 2]);
    }
    long long maxpairs = 0;
    std::vector<unsigned> toff(natoms);
    long long tstride = 0;
    for (int64_t a = 0; a < natoms; ++a) {
        if (type_id[a] < 0 || type_id[a] >= nt) return fail(MDSF_EINVAL, "type_id[%lld]=%d out of range", (long long)a, type_id[a]);
        maxpairs += bound[type_id[a]];
        toff[a] = (unsigned)tstride;
        const int* hw = &h->halfw_host[type_id[a] * 3];
        tstride += 2 * (hw[0] + hw[1] + hw[2]);
    }
    if (tstride * h->F >= (1LL << 32)) return fail(MDSF_EINVAL, "factor tables overflow 32-bit offsets; lower batch_frames");
    h->gp.tstride = tstride;
    // Nz = R*R: the z pass runs as two radix-R stages whose (transposed) output is in natural frequency order
    h->zfast = 0;
    if (h->native_fft && h->ax[2].plan.nstages == 2 && h->ax[2].plan.radix[0] == h->ax[2].plan.radix[1] &&
        (h->ax[2].plan.radix[0] == 16 || h->ax[2].plan.radix[0] == 8) && !getenv("MDSF_NO_ZFAST")) {
        h->zfast = h->ax[2].plan.radix[0];
        std::vector<int> ident(g0.n[2]);
        for (int k = 0; k < g0.n[2]; ++k) ident[k] = k;
        CU(cudaMemcpy(h->ax[2].d_rev, ident.data(), sizeof(int) * g0.n[2], cudaMemcpyHostToDevice));
    }
    // ---- splat mode: small stamps -> fixed-point scatter into L2-resident slabs; large stamps -> owner tiles
    {
        double terms = 0, amax = 0;
        for (int64_t a = 0; a < natoms; ++a) {
            const int* hw = &h->halfw_host[type_id[a] * 3];
            terms += 8.0 * hw[0] * hw[1] * hw[2];
        }
        // measured on B200: the shared-memory fixed-point tile mode wins for small and
        // mid-size stamps (c2: 64 cells, c1: ~800 cells per atom); scatter and owner stay selectable
        h->scatter = h->want_mode == 2;
        if (h->scatter) h->tile_atomic = false;
        if (h->scatter) {
            if ((long long)natoms * h->F >= (1LL << MDSF_ENTRY_BITS)) return fail(MDSF_EINVAL, "natoms*batch_frames exceeds 2^30 in scatter mode");
            std::vector<double> amp(nt);
            CU(cudaMemcpy(amp.data(), h->d_amp, sizeof(double) * nt, cudaMemcpyDeviceToHost));
            for (int t = 0; t < nt; ++t) amax = std::max(amax, amp[t]);
            int e = 0;
            (void)std::frexp(amax, &e);                   // amax < 2^e
            h->sp.scale = std::ldexp(1.0, 52 - e);
            h->sp.inv_scale = std::ldexp(1.0, e - 52);
            const int npairs = h->F / 2;
            const double per_plane = (double)npairs * g0.n[1] * g0.n[2] * 16.0;
            // R slab accumulators form a ring that must stay resident in L2 (measured on B200: a 17 MB ring
            // stays resident, a 50 MB ring does not) -> default 8.4 MB per slab, ring of 3
            const double budget = (getenv("MDSF_SLAB_MB") ? atof(getenv("MDSF_SLAB_MB")) : 8.5) * 1048576.0;
            h->ring = getenv("MDSF_SLAB_RING") ? std::max(2, atoi(getenv("MDSF_SLAB_RING"))) : 3;
            int X = (int)std::floor(budget / per_plane);
            X = std::max(1, std::min(X, std::min(g0.n[0], 1023)));
            h->sp.X = X;
            h->sp.nslabs = (g0.n[0] + X - 1) / X;
            if (h->sp.nslabs > 4096) return fail(MDSF_EINVAL, "too many slabs");
            long long cap = 0;
            std::vector<long long> sb(nt, 1);
            for (int t = 0; t < nt; ++t) {
                const int A = h->halfw_host[t * 3], N = g0.n[0];
                for (int ir = A - g0.nb; ir <= N + g0.nb - A; ++ir) sb[t] = std::max<long long>(sb[t], stamp_tiles_1d(ir, A, N, X));
            }
            for (int64_t a = 0; a < natoms; ++a) cap += sb[type_id[a]];
            h->entries_cap = cap * h->F;
            if (h->entries_cap >= (1LL << 32) - 2) return fail(MDSF_EINVAL, "slab entry capacity overflows 32 bits");
            const size_t acc_cells = (size_t)npairs * X * g0.n[1] * g0.n[2];
            h->acc_cells = acc_cells;
            {   // keep the accumulator slabs resident in L2: persisting window on the compute stream
                cudaDeviceProp prop;
                CU(cudaGetDeviceProperties(&prop, h->device));
                const size_t want = acc_cells * 16 * h->ring;
                const size_t carve = std::min<size_t>(want, (size_t)prop.persistingL2CacheMaxSize);
                if (carve > 0 && getenv("MDSF_L2_PERSIST")) {      // opt-in: measured slower on B200 (it starves the y/x passes of L2)
                    CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
                    h->l2_window = std::min<size_t>(want, (size_t)prop.accessPolicyMaxWindowSize);
                    h->l2_ratio = (float)std::min(1.0, (double)carve / (double)h->l2_window);
                }
            }
            CU(cudaMalloc(&h->d_acc, acc_cells * 16 * h->ring));    // ring of slab accumulators
            CU(cudaMemset(h->d_acc, 0, acc_cells * 16 * h->ring));
            CU(cudaMalloc(&h->d_step_start, sizeof(unsigned) * (h->sp.nslabs + 2)));
            CU(cudaMalloc(&h->d_ctl, sizeof(unsigned) * (2 + 2 * h->sp.nslabs)));
            if (h->l2_window) {
                cudaStreamAttrValue attr{};
                attr.accessPolicyWindow.base_ptr = h->d_acc;
                attr.accessPolicyWindow.num_bytes = h->l2_window;
                attr.accessPolicyWindow.hitRatio = h->l2_ratio;
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                CU(cudaStreamSetAttribute(h->s_comp, cudaStreamAttributeAccessPolicyWindow, &attr));
            }
            CU(cudaMalloc(&h->d_slab_count, sizeof(unsigned) * (h->sp.nslabs + 1)));
            CU(cudaMalloc(&h->d_slab_start, sizeof(unsigned) * (h->sp.nslabs + 1)));
            CU(cudaMalloc(&h->d_slab_cursor, sizeof(unsigned) * (h->sp.nslabs + 1)));
            CU(cudaMalloc(&h->d_entries, sizeof(unsigned) * std::max(1LL, h->entries_cap)));
            h->zcol = MDSF_PIPE_THREADS / 16;
            while (h->zcol > 1 && (size_t)2 * h->zcol * g0.nzp * 8 > 80 * 1024) h->zcol >>= 1;
            int per_sm = 0;
            const size_t zsm = (size_t)2 * h->zcol * g0.nzp * 8 + (size_t)2 * g0.n[2] * 8;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, slab_pipeline_kernel, MDSF_PIPE_THREADS, zsm));
            if (per_sm < 1) return fail(MDSF_EINVAL, "slab pipeline kernel does not fit on an SM");
            h->pipe_grid = per_sm * h->nsm;
        }
    }
    h->natoms = natoms;
    h->gp.natoms = (int)natoms;
    h->maxpairs_frame = maxpairs;
    const long long cap = maxpairs * h->F;
    if (cap >= (1LL << 32) - 2) return fail(MDSF_EINVAL, "pair capacity %lld overflows 32-bit offsets; lower batch_frames", cap);
    const long long nkeys = (long long)h->F * g0.ntx * g0.nty * g0.nslab;
    if (nkeys >= (1LL << 31)) return fail(MDSF_EINVAL, "too many tiles per batch");
    h->sort_bits = 1;
    while ((1LL << h->sort_bits) <= nkeys) ++h->sort_bits;

    // splat shared-memory budget -> pairs per chunk; per-pair table stride 2^logS >= tx + ty + max 2Az
    // stamps taller than `zstage` cells (rare heavy ions next to light atoms) keep EZ in global memory, so they do
    // not inflate every pair's slot
    h->zstage = zmax;
    if (zmax > 16) {
        long long tall = 0;
        for (int64_t a = 0; a < natoms; ++a) tall += 2 * h->halfw_host[type_id[a] * 3 + 2] > 16;
        if (tall * 4 < natoms) h->zstage = 16;
    }
    if (getenv("MDSF_ZSTAGE")) h->zstage = atoi(getenv("MDSF_ZSTAGE"));
    h->ez_global = zmax > h->zstage;
    const int zslot = std::max(1, std::min(zmax, h->zstage));
    h->logS = 0;
    while ((1 << h->logS) < g0.tx + g0.ty + zslot) ++h->logS;
    const size_t tile_b = (size_t)2 * g0.tx * g0.ty * g0.nzp * 8;
    int chunk = 128;
    // tables (also hold r in the general-ucell path and the z twiddles after the splat), pair info, hit masks
    auto smem_for = [&](int c) { return tile_b + std::max((size_t)2 * c * 8 << h->logS, (size_t)2 * g0.n[2] * 8) + (size_t)2 * c * sizeof(PairInfo) + 2 * 4 * 32 * 4 + (size_t)2 * (c + 8) * 4 + (size_t)2 * c * g0.tx * g0.ty + 64; };
    const size_t soft = tile_b <= 80 * 1024 ? 113 * 1024 : kMaxSmem;    // two CTAs per SM when the tile allows
    // Nz = 256 tile mode: the 16x16 split's inter-stage twiddles get their own 4 KB, filled by cp.async at kernel
    // start, instead of a global -> shared copy between the splat and the FFT (one exposed round trip less per CTA)
    const bool tw_pref = h->tile_atomic && h->native_fft && h->zfast == 16 && h->ax[2].d_tw16 && !h->cfg.keep_density &&
                         (!getenv("MDSF_TW_PREFETCH") || atoi(getenv("MDSF_TW_PREFETCH")) != 0);
    const size_t tw_extra = tw_pref ? 4096 + 16 : 0;
    while (chunk > 32 && smem_for(chunk) + tw_extra > soft) chunk -= 32;
    if (smem_for(chunk) + tw_extra > (size_t)kMaxSmem) return fail(MDSF_EINVAL, "splat tile does not fit shared memory (%zu bytes)", smem_for(chunk));
    h->chunk = chunk;
    h->tw16_off = tw_pref ? (int)((smem_for(chunk) + 15) / 16 * 16) : 0;
    h->splat_smem = smem_for(chunk) + tw_extra;

    CU(cudaMalloc(&h->d_type, sizeof(int) * natoms));
    CU(cudaMemcpy(h->d_type, type_id, sizeof(int) * natoms, cudaMemcpyHostToDevice));
    for (int s = 0; s < kSlots; ++s) CU(cudaMalloc(&h->d_stage[s], h->csize * 3 * natoms * h->F));
    CU(cudaMalloc(&h->d_toff, sizeof(unsigned) * natoms));
    CU(cudaMemcpy(h->d_toff, toff.data(), sizeof(unsigned) * natoms, cudaMemcpyHostToDevice));
    h->tt.toff = h->d_toff;
    h->nsets = (h->scatter || getenv("MDSF_ONE_STREAM")) ? 1 : 2;
    if (getenv("MDSF_PF_DIST")) h->pf_dist = std::max(0, atoi(getenv("MDSF_PF_DIST")));
    h->pair_records = !getenv("MDSF_PAIR_RECORDS") || atoi(getenv("MDSF_PAIR_RECORDS")) != 0;
    h->direct_bin = h->tile_atomic && !h->scatter && (!getenv("MDSF_DIRECT_BIN") || atoi(getenv("MDSF_DIRECT_BIN")) != 0);
    {
        size_t b1 = 0, b2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b1, (unsigned*)nullptr, (unsigned*)nullptr, (long long)natoms * h->F, h->s_comp);
        cub::DeviceRadixSort::SortPairs(nullptr, b2, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, cap, 0, h->sort_bits, h->s_comp);
        size_t b3 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b3, (unsigned*)nullptr, (unsigned*)nullptr, (long long)nkeys + 2, h->s_comp);
        h->cub_bytes = std::max(std::max(b1, b2), b3) + 256;
    }
    for (int p = 0; p < h->nsets; ++p) {
        mdsf_handle::PrepSet& ps = h->sets[p];
        CU(cudaMalloc(&ps.recs, sizeof(AtomRec) * natoms * h->F));
        CU(cudaMalloc(&ps.tables, sizeof(double) * (std::max(1LL, tstride * h->F) + 32)));     // + slack: the splat prefetches one line past a table's start
        CU(cudaMalloc(&ps.cnt, sizeof(unsigned) * natoms * h->F));
        CU(cudaMalloc(&ps.off, sizeof(unsigned) * natoms * h->F));
        if (!h->scatter) {
            for (int i = 0; i < 2; ++i) {
                CU(cudaMalloc(&ps.keys[i], sizeof(unsigned) * std::max(1LL, cap)));
                CU(cudaMalloc(&ps.vals[i], sizeof(unsigned) * std::max(1LL, cap)));
            }
            CU(cudaMalloc(&ps.tile_start, sizeof(unsigned) * (nkeys + 2)));
            CU(cudaMalloc(&ps.counter, sizeof(unsigned) * 2 * (nkeys + 2)));
            if (h->direct_bin && h->pair_records) CU(cudaMalloc(&ps.prec, sizeof(uint4) * std::max(1LL, cap)));
            CU(cudaMalloc(&ps.cub, h->cub_bytes));
        }
        CU(cudaEventCreateWithFlags(&ps.ev_binned, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ps.ev_consumed, cudaEventDisableTiming));
    }
    h->d_recs = h->sets[0].recs; h->d_tables = h->sets[0].tables; h->d_cnt = h->sets[0].cnt; h->d_off = h->sets[0].off;
    return MDSF_OK;
}

extern "C" int mdsf_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(MDSF_EINVAL, "null argument");
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return MDSF_OK;
}
extern "C" int mdsf_host_free(void* p) { if (p) CU(cudaFreeHost(p)); return MDSF_OK; }
extern "C" int mdsf_host_register(void* p, size_t bytes) { CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault)); return MDSF_OK; }
extern "C" int mdsf_host_unregister(void* p) { CU(cudaHostUnregister(p)); return MDSF_OK; }

// ------------------------------------------------------------------------------------------
// FFT + accumulation of the pair volumes currently in d_vol (nf frames -> (nf+1)/2 pairs).
// 1024-point axes (8*8*16) through the three-stage register kernels.  Measured on c5 (2 frames per step): y pass
// 16.0 -> 15.2 ms, x pass 19.8 -> 25.2 ms (one 256-thread CTA per SM against the generic kernel's 512 threads), so
// the y pass takes them by default and the x pass does not; MDSF_FFT3_1024_Y / MDSF_FFT3_1024_X = 0/1 for A/B runs.
static bool fft3_1024_enabled(bool xpass) {
    const char* e = getenv(xpass ? "MDSF_FFT3_1024_X" : "MDSF_FFT3_1024_Y");
    return e ? atoi(e) != 0 : !xpass;
}

// `z_done`: the z pass already ran inside the fused splat kernel.
static int transform_and_accumulate(mdsf_handle* h, int nf, bool z_done, cudaEvent_t* tv = nullptr) {
    const GridParams& gp = h->gp;
    const int npairs = (nf + 1) / 2;
    if (h->native_fft) {
        if (!z_done) {
            const long long ncolumns = (long long)gp.n[0] * gp.n[1];
            int ncol = 16;
            while (ncol > 1 && (size_t)2 * ncol * gp.nzp * 8 > 96 * 1024) ncol >>= 1;
            const size_t sm = (size_t)2 * ncol * gp.nzp * 8 + (size_t)2 * gp.n[2] * 8;
            dim3 grid((unsigned)((ncolumns + ncol - 1) / ncol), npairs);
            fft_z_kernel<<<grid, 256, sm, h->s_comp>>>(h->d_vol, h->ax[2].plan, h->ax[2].d_tw, ncolumns, ncol, gp.nzp, gp.pad_shift, h->zfast);
            ++h->launches;
        }
        if (tv && !h->overlap) CU(cudaEventRecord(tv[3], h->s_comp));
        if (tv) CU(cudaEventRecord(tv[7], h->s_comp));
        {
            const size_t sm = (size_t)2 * gp.n[1] * h->Wy * 8 + (size_t)2 * gp.n[1] * 8;
            dim3 grid((gp.n[2] + h->Wy - 1) / h->Wy, gp.n[0], npairs);
            int logw = 0; while ((1 << logw) < h->Wy) ++logw;
            const FftPlan& yp = h->ax[1].plan;
            const bool fast = yp.nstages == 2 && yp.radix[0] == yp.radix[1] && (gp.n[1] / yp.radix[0]) * h->Wy == h->thr_y && h->thr_y == MDSF_PASS_THREADS &&
                              (yp.radix[0] == 16 || yp.radix[0] == 8) && !getenv("MDSF_NO_YFAST");
            const size_t sya = sm + (size_t)yp.radix[0] * h->thr_y * 16;
            dim3 grid_a((gp.n[2] + h->Wy - 1) / h->Wy, gp.n[0]);
            const bool three = yp.nstages == 3 && h->Wy == 8 && !getenv("MDSF_NO_FFT3");
            const bool three8 = three && yp.radix[0] == 8 && yp.radix[1] == 8 && yp.radix[2] == 8;          // 512
            const bool three768 = three && yp.radix[0] == 16 && yp.radix[1] == 16 && yp.radix[2] == 3;     // 768
            const bool three1024 = three && yp.radix[0] == 8 && yp.radix[1] == 8 && yp.radix[2] == 16 && fft3_1024_enabled(false);
            if (three8)
                fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_Y, false><<<grid, 256, sm, h->s_comp>>>(h->d_vol, nullptr, h->ax[1].d_tw, gp.n[0], gp.n[2], npairs);
            else if (three768)
                fft3_pass_kernel<16, 16, 3, 3, 256, 2, false><<<grid, 256, sm, h->s_comp>>>(h->d_vol, nullptr, h->ax[1].d_tw, gp.n[0], gp.n[2], npairs);
            else if (three1024)
                fft3_pass_kernel<8, 8, 16, 3, 256, 1, false><<<grid, 256, sm, h->s_comp>>>(h->d_vol, nullptr, h->ax[1].d_tw, gp.n[0], gp.n[2], npairs);
            else if (fast && h->y_async && yp.radix[0] == 16)
                fft_y_async_kernel<16, 16><<<grid_a, h->thr_y, sya, h->s_comp>>>(h->d_vol, h->ax[1].d_tw, gp.n[0], gp.n[2], logw, npairs);
            else if (fast && h->y_async)
                fft_y_async_kernel<8, 8><<<grid_a, h->thr_y, sya, h->s_comp>>>(h->d_vol, h->ax[1].d_tw, gp.n[0], gp.n[2], logw, npairs);
            else if (fast && yp.radix[0] == 16)
                fft_y_fast_kernel<16, 16><<<grid, h->thr_y, sm, h->s_comp>>>(h->d_vol, h->ax[1].d_tw, gp.n[0], gp.n[2], logw);
            else if (fast)
                fft_y_fast_kernel<8, 8><<<grid, h->thr_y, sm, h->s_comp>>>(h->d_vol, h->ax[1].d_tw, gp.n[0], gp.n[2], logw);
            else if (h->thr_y == 512)
                fft_y_kernel<512, 1><<<grid, 512, sm, h->s_comp>>>(h->d_vol, h->ax[1].plan, h->ax[1].d_tw, gp.n[0], gp.n[1], gp.n[2], h->Wy, logw);
            else if (h->thr_y == 256)
                fft_y_kernel<256, 2><<<grid, 256, sm, h->s_comp>>>(h->d_vol, h->ax[1].plan, h->ax[1].d_tw, gp.n[0], gp.n[1], gp.n[2], h->Wy, logw);
            else
                fft_y_kernel<128, MDSF_PASS_MINBLOCKS><<<grid, 128, sm, h->s_comp>>>(h->d_vol, h->ax[1].plan, h->ax[1].d_tw, gp.n[0], gp.n[1], gp.n[2], h->Wy, logw);
            ++h->launches;
        }
        if (tv) CU(cudaEventRecord(tv[4], h->s_comp));
        {
            const size_t sm = (size_t)3 * gp.n[0] * h->Wx * 8 + (size_t)2 * gp.n[0] * 8;
            int logw = 0; while ((1 << logw) < h->Wx) ++logw;
            dim3 grid((gp.n[2] + h->Wx - 1) / h->Wx, gp.n[1]);
            const FftPlan& xp = h->ax[0].plan;
            const bool fast = xp.nstages == 2 && xp.radix[0] == xp.radix[1] && (gp.n[0] / xp.radix[0]) * h->Wx == h->thr_x && h->thr_x == MDSF_PASS_THREADS &&
                              (xp.radix[0] == 16 || xp.radix[0] == 8) && !getenv("MDSF_NO_XFAST");
            const size_t smf = (size_t)2 * gp.n[0] * h->Wx * 8 + (size_t)2 * gp.n[0] * 8;
            const size_t sma = smf + (size_t)xp.radix[0] * h->thr_x * 16;     // + cp.async staging slots
            const bool three = xp.nstages == 3 && h->Wx == 8 && !getenv("MDSF_NO_FFT3");
            const bool three8 = three && xp.radix[0] == 8 && xp.radix[1] == 8 && xp.radix[2] == 8;
            const bool three768 = three && xp.radix[0] == 16 && xp.radix[1] == 16 && xp.radix[2] == 3;
            const bool three1024 = three && xp.radix[0] == 8 && xp.radix[1] == 8 && xp.radix[2] == 16 && fft3_1024_enabled(true);
            if (three8)
                fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_X, true><<<grid, 256, smf, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].d_tw, gp.n[1], gp.n[2], npairs);
            else if (three768)
                fft3_pass_kernel<16, 16, 3, 3, 256, 2, true><<<grid, 256, smf, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].d_tw, gp.n[1], gp.n[2], npairs);
            else if (three1024)
                fft3_pass_kernel<8, 8, 16, 3, 256, 1, true><<<grid, 256, smf, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].d_tw, gp.n[1], gp.n[2], npairs);
            else if (fast && h->x_async && xp.radix[0] == 16)
                fft_x_accum_async_kernel<16, 16><<<grid, h->thr_x, sma, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].d_tw, gp.n[1], gp.n[2], logw, npairs);
            else if (fast && h->x_async)
                fft_x_accum_async_kernel<8, 8><<<grid, h->thr_x, sma, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].d_tw, gp.n[1], gp.n[2], logw, npairs);
            else if (fast && xp.radix[0] == 16)
                fft_x_accum_fast_kernel<16, 16><<<grid, h->thr_x, smf, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].d_tw, gp.n[1], gp.n[2], logw, npairs);
            else if (fast)
                fft_x_accum_fast_kernel<8, 8><<<grid, h->thr_x, smf, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].d_tw, gp.n[1], gp.n[2], logw, npairs);
            else if (h->thr_x == 512)
                fft_x_accum_kernel<512, 1><<<grid, 512, sm, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].plan, h->ax[0].d_tw,
                                                                         gp.n[0], gp.n[1], gp.n[2], h->Wx, logw, npairs);
            else if (h->thr_x == 256)
                fft_x_accum_kernel<256, 2><<<grid, 256, sm, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].plan, h->ax[0].d_tw,
                                                                         gp.n[0], gp.n[1], gp.n[2], h->Wx, logw, npairs);
            else
                fft_x_accum_kernel<128, MDSF_PASS_MINBLOCKS><<<grid, 128, sm, h->s_comp>>>(h->d_vol, h->d_P, h->ax[0].plan, h->ax[0].d_tw,
                                                                                           gp.n[0], gp.n[1], gp.n[2], h->Wx, logw, npairs);
            ++h->launches;
        }
    } else {
        if (tv) { CU(cudaEventRecord(tv[3], h->s_comp)); CU(cudaEventRecord(tv[7], h->s_comp)); }
        if (npairs == h->cufft_batch) {
            CF(cufftExecZ2Z(h->cufft_plan, (cufftDoubleComplex*)h->d_vol, (cufftDoubleComplex*)h->d_vol, CUFFT_FORWARD));
        } else {   // partial last batch: zero the unused pair volumes and transform the whole batch
            CU(cudaMemsetAsync(h->d_vol + (long long)npairs * h->ncell, 0, sizeof(double2) * h->ncell * (h->cufft_batch - npairs), h->s_comp));
            CF(cufftExecZ2Z(h->cufft_plan, (cufftDoubleComplex*)h->d_vol, (cufftDoubleComplex*)h->d_vol, CUFFT_FORWARD));
        }
        if (tv) CU(cudaEventRecord(tv[4], h->s_comp));
        accumulate_power_kernel<<<grid_for(h->ncell, 256, h->nsm), 256, 0, h->s_comp>>>(h->d_vol, h->d_P, h->ncell, npairs);
        ++h->launches;
    }
    CU(cudaGetLastError());
    return MDSF_OK;
}

template <typename C, typename P>
static void launch_prep(mdsf_handle* h, cudaStream_t st, void* stage, const BatchScales& sc, int nf, long long wlo, long long whi) {
    const long long total = (long long)nf * h->natoms;
    prep_atoms_kernel<C, P><<<grid_for(total, 256, h->nsm), 256, 0, st>>>(
        (C*)stage, h->d_type, h->d_recs, h->d_cnt, h->d_tables, h->gp, h->tt, sc, nf, wlo, whi, h->d_err, h->direct_bin ? h->d_counter : nullptr,
        h->mono, h->mono_sin, h->mono_cos);
}

static int run_batch(mdsf_handle* h, char* src, int nf, const double* scale, long long wlo, long long whi, int write_back) {
    const GridParams& gp = h->gp;
    const int slot = h->next_slot;
    h->next_slot = (h->next_slot + 1) % kSlots;
    const size_t bytes = h->csize * 3 * h->natoms * nf;
    // H2D on the copy stream once the slot's previous consumers are done
    if (h->slot_used[slot]) {
        CU(cudaStreamWaitEvent(h->s_copy, h->ev_free[slot], 0));
        CU(cudaStreamWaitEvent(h->s_copy, h->ev_back[slot], 0));
    }
    // prep/bin outputs of this batch go to set p; with two sets they are produced on s_prep while the
    // previous batch is still in its splat/FFT kernels on s_comp
    const int p = (int)(h->batch_counter++ % h->nsets);
    mdsf_handle::PrepSet& ps = h->sets[p];
    // overlap mode: prep/bin and the splat share the in-order stream s_splat; the passes keep s_comp
    cudaStream_t sp = h->overlap ? h->s_splat : (h->nsets == 2 ? h->s_prep : h->s_comp);
    cudaStream_t ss = h->overlap ? h->s_splat : h->s_comp;        // stream of the splat kernel
    const int v = h->overlap ? (int)((h->batch_counter - 1) & 1) : 0;
    h->d_vol = h->d_volset[v];
    h->d_recs = ps.recs; h->d_cnt = ps.cnt; h->d_off = ps.off; h->d_tables = ps.tables;
    h->d_keys[0] = ps.keys[0]; h->d_keys[1] = ps.keys[1]; h->d_vals[0] = ps.vals[0]; h->d_vals[1] = ps.vals[1];
    h->d_tile_start = ps.tile_start; h->d_cub = ps.cub;
    cudaEvent_t* tv = nullptr;
    if (h->timing) {
        for (int i = 0; i < 8; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); h->tev.push_back(e); }
        tv = &h->tev[h->tev.size() - 8];
        CU(cudaEventRecord(tv[0], h->s_copy));
    }
    CU(cudaMemcpyAsync(h->d_stage[slot], src, bytes, cudaMemcpyDefault, h->s_copy));
    CU(cudaEventRecord(h->ev_h2d[slot], h->s_copy));
    CU(cudaStreamWaitEvent(sp, h->ev_h2d[slot], 0));
    if (ps.used && h->nsets == 2 && !h->overlap) CU(cudaStreamWaitEvent(sp, ps.ev_consumed, 0));     // the splat of batch b-2 has read this set
    // start after the splat of batch b-1: prep+bin then overlap its HBM-bound y/x passes instead of fighting the
    // issue-bound splat kernel for the SMs
    if (h->nsets == 2 && !h->overlap && h->sets[1 - p].used && !getenv("MDSF_PREP_EARLY")) CU(cudaStreamWaitEvent(sp, h->sets[1 - p].ev_consumed, 0));
    if (tv) CU(cudaEventRecord(tv[1], sp));

    BatchScales sc;
    for (int f = 0; f < nf; ++f) for (int d = 0; d < 3; ++d) sc.a[f][d] = scale[f * 3 + d];
    h->d_counter = ps.counter;
    if (h->direct_bin) {      // list lengths are counted by K1 itself; [nkeys+1] lengths, then [nkeys+1] cursors
        const unsigned nk = (unsigned)((nf + (nf & 1)) * gp.ntx * gp.nty * gp.nslab);
        CU(cudaMemsetAsync(ps.counter, 0, sizeof(unsigned) * 2 * (nk + 1), sp));
    }
    const bool c32 = h->cfg.coord_dtype == MDSF_F32, p32 = h->cfg.arith_dtype == MDSF_F32;
    if (c32 && p32) launch_prep<float, float>(h, sp, h->d_stage[slot], sc, nf, wlo, whi);
    else if (c32) launch_prep<float, double>(h, sp, h->d_stage[slot], sc, nf, wlo, whi);
    else launch_prep<double, double>(h, sp, h->d_stage[slot], sc, nf, wlo, whi);
    ++h->launches;
    CU(cudaEventRecord(h->ev_prep[slot], sp));
    CU(cudaEventRecord(h->ev_free[slot], sp));          // the staging slot is only read (and rewritten) by K1
    if (write_back) {
        CU(cudaStreamWaitEvent(h->s_back, h->ev_prep[slot], 0));
        CU(cudaMemcpyAsync(src, h->d_stage[slot], bytes, cudaMemcpyDefault, h->s_back));
    }
    CU(cudaEventRecord(h->ev_back[slot], h->s_back));

    const int npairs = (nf + 1) / 2;
    if (h->scatter) {
        // K2s: counting sort of atom images by x slab (order inside a bin is irrelevant: integer adds commute)
        const SlabParams& sp = h->sp;
        CU(cudaMemsetAsync(h->d_slab_count, 0, sizeof(unsigned) * (sp.nslabs + 1), h->s_comp));
        bin_slabs_kernel<0><<<h->nsm * 4, 256, sizeof(unsigned) * 2 * sp.nslabs, h->s_comp>>>(h->d_recs, h->d_cnt, h->d_slab_count, nullptr, gp, h->tt, sp, nf);
        scan_slabs_kernel<<<1, 1024, 0, h->s_comp>>>(h->d_slab_count, h->d_slab_start, h->d_slab_cursor, sp.nslabs, h->d_step_start,
                                                     MDSF_SC_ENTRIES, sp.X, gp.n[0], gp.n[1], h->zcol, npairs);
        CU(cudaMemsetAsync(h->d_ctl, 0, sizeof(unsigned) * (2 + 2 * sp.nslabs), h->s_comp));
        bin_slabs_kernel<1><<<h->nsm * 4, 256, sizeof(unsigned) * 2 * sp.nslabs, h->s_comp>>>(h->d_recs, h->d_cnt, h->d_slab_cursor, h->d_entries, gp, h->tt, sp, nf);
        h->launches += 3;
        if (tv) { CU(cudaEventRecord(tv[2], h->s_comp)); CU(cudaEventRecord(tv[6], h->s_comp)); }
        // K3s/K3z: persistent cooperative kernel, scatter(s) overlapped with the z pass of slab s-1
        FftPlan zplan = h->native_fft ? h->ax[2].plan : FftPlan{gp.n[2], 0, {0}};
        size_t zsm = (size_t)2 * h->zcol * gp.nzp * 8 + (size_t)2 * gp.n[2] * 8;
        const double2* twz = h->ax[2].d_tw;
        GridParams gpl = gp;
        TypeTable ttl = h->tt;
        SlabParams spl = sp;
        int np = npairs, ncol = h->zcol;
        size_t cells = h->acc_cells;
        int ring = h->ring, zfast = h->zfast;
        void* args[] = {&h->d_recs, &h->d_entries, &h->d_slab_start, &h->d_step_start, &h->d_ctl, &h->d_tables, &h->d_acc, &cells, &ring,
                        &h->d_vol, &h->d_dump, &zplan, &twz, &gpl, &ttl, &spl, &np, &ncol, &zfast, &h->d_err};
        CU(cudaLaunchCooperativeKernel((void*)slab_pipeline_kernel, dim3(h->pipe_grid), dim3(MDSF_PIPE_THREADS), args, zsm, h->s_comp));
        ++h->launches;
        CU(cudaGetLastError());
    } else {
    // deterministic binning: scan -> emit -> stable radix sort by (frame, tile) -> list starts
    const long long total = (long long)nf * h->natoms;
    const long long cap = h->maxpairs_frame * nf;
    // an odd batch gets a phantom last frame with empty lists (imaginary part of the last pair)
    const unsigned nkeys = (unsigned)((nf + (nf & 1)) * gp.ntx * gp.nty * gp.nslab);
    size_t cb = h->cub_bytes;
    if (h->direct_bin) {
        // tile mode: order inside a list is irrelevant (integer adds commute) -> counting sort with atomics
        cub::DeviceScan::ExclusiveSum(h->d_cub, cb, ps.counter, h->d_tile_start, (long long)nkeys + 1, sp);
        bin_pairs_kernel<true><<<grid_for(total, 256, h->nsm), 256, 0, sp>>>(h->d_recs, h->d_cnt, ps.counter + nkeys + 1, h->d_tile_start, h->d_vals[1], ps.prec, gp, h->tt, nf);
        h->launches += 1;
    } else {
    cub::DeviceScan::ExclusiveSum(h->d_cub, cb, h->d_cnt, h->d_off, total, sp);
    fill_u32_kernel<<<grid_for(cap, 256, h->nsm), 256, 0, sp>>>(h->d_keys[0], nkeys, cap);
    emit_pairs_kernel<<<grid_for(total, 256, h->nsm), 256, 0, sp>>>(h->d_recs, h->d_cnt, h->d_off, h->d_keys[0], h->d_vals[0], gp, h->tt, nf);
    int bits = 1;
    while ((1LL << bits) <= (long long)nkeys) ++bits;
    cb = h->cub_bytes;
    cub::DeviceRadixSort::SortPairs(h->d_cub, cb, h->d_keys[0], h->d_keys[1], h->d_vals[0], h->d_vals[1], cap, 0, bits, sp);
    tile_starts_kernel<<<grid_for(cap + 1, 256, h->nsm), 256, 0, sp>>>(h->d_keys[1], cap, nkeys, h->d_tile_start);
    h->launches += 3;
    }
    if (tv) CU(cudaEventRecord(tv[2], sp));
    if (h->nsets == 2 && !h->overlap) {
        CU(cudaEventRecord(ps.ev_binned, sp));
        CU(cudaStreamWaitEvent(h->s_comp, ps.ev_binned, 0));
    }
    if (h->overlap && h->vol_used[v]) CU(cudaStreamWaitEvent(ss, h->ev_volfree[v], 0));   // the x pass of batch b-2 has read this volume set
    if (tv) CU(cudaEventRecord(tv[6], ss));

    // splat (+ fused z FFT on the native path)
    dim3 grid(gp.ntx * gp.nty, npairs);
    const bool use_prec = h->direct_bin && ps.prec != nullptr;
    const unsigned* list = use_prec ? reinterpret_cast<const unsigned*>(ps.prec) : h->d_vals[1];
#define MDSF_SPLAT_LAUNCH(FUSE, ATOM, EZG, PREC) MDSF_SPLAT_LAUNCH5(FUSE, ATOM, EZG, PREC, false)
#define MDSF_SPLAT_LAUNCH5(FUSE, ATOM, EZG, PREC, T44)                                                                    \
    splat_zfft_kernel<FUSE, ATOM, EZG, PREC, T44><<<grid, 256, h->splat_smem, ss>>>(h->d_recs, list, h->d_tile_start, h->d_vol, \
        h->d_dump, gp, h->tt, h->ax[2].plan, h->ax[2].d_tw, h->d_tables, h->chunk, h->logS, h->zfast, h->zstage, h->d_err, h->ax[2].d_tw16, h->tw16_off, h->pf_dist)
    switch ((use_prec ? 8 : 0) | (h->native_fft ? 4 : 0) | (h->tile_atomic ? 2 : 0) | (h->ez_global ? 1 : 0)) {
        case 0: MDSF_SPLAT_LAUNCH(false, false, false, false); break;
        case 1: MDSF_SPLAT_LAUNCH(false, false, true, false); break;
        case 2: MDSF_SPLAT_LAUNCH(false, true, false, false); break;
        case 3: MDSF_SPLAT_LAUNCH(false, true, true, false); break;
        case 4: MDSF_SPLAT_LAUNCH(true, false, false, false); break;
        case 5: MDSF_SPLAT_LAUNCH(true, false, true, false); break;
        case 6: MDSF_SPLAT_LAUNCH(true, true, false, false); break;
        case 7: MDSF_SPLAT_LAUNCH(true, true, true, false); break;
        case 10: MDSF_SPLAT_LAUNCH(false, true, false, true); break;
        case 11: MDSF_SPLAT_LAUNCH(false, true, true, true); break;
        case 14:
            if (gp.tx == 4 && gp.ty == 4 && h->logS == 4 && !getenv("MDSF_NO_T44"))
                MDSF_SPLAT_LAUNCH5(true, true, false, true, true);
            else MDSF_SPLAT_LAUNCH(true, true, false, true);
            break;
        case 15: MDSF_SPLAT_LAUNCH(true, true, true, true); break;
        default: return fail(MDSF_ESTATE, "pair records without tile mode");
    }
    ++h->launches;
    CU(cudaGetLastError());
    if (h->nsets == 2 && !h->overlap) CU(cudaEventRecord(ps.ev_consumed, h->s_comp));
    if (h->overlap) {
        if (tv) CU(cudaEventRecord(tv[3], ss));
        CU(cudaEventRecord(h->ev_splat[v], ss));
        CU(cudaStreamWaitEvent(h->s_comp, h->ev_splat[v], 0));
    }
    }
    ps.used = true;
    int rc = transform_and_accumulate(h, nf, h->native_fft, tv);   // z pass already done on the native path
    if (rc) return rc;
    if (h->overlap) { CU(cudaEventRecord(h->ev_volfree[v], h->s_comp)); h->vol_used[v] = true; }
    if (tv) { CU(cudaEventRecord(tv[5], h->s_comp)); ++h->timed_batches; }
    h->slot_used[slot] = true;
    h->frames_done += nf;
    h->last_batch_frames = nf;
    return MDSF_OK;
}

extern "C" int mdsf_push_frames(mdsf_handle* h, void* coords, int64_t nframes, const double* scale,
                                int64_t wrap_lo, int64_t wrap_hi, int32_t write_back) {
    if (!h || !coords || !scale) return fail(MDSF_EINVAL, "null argument");
    if (!h->d_type) return fail(MDSF_ESTATE, "mdsf_set_atoms must be called before mdsf_push_frames");
    if (nframes < 0) return fail(MDSF_EINVAL, "negative frame count");
    CU(cudaSetDevice(h->device));
    const size_t frame_bytes = h->csize * 3 * h->natoms;
    int64_t done = 0;
    while (done < nframes) {
        const int nf = (int)std::min<int64_t>(h->F, nframes - done);
        int rc = run_batch(h, (char*)coords + frame_bytes * done, nf, scale + 3 * done, wrap_lo, wrap_hi, write_back);
        if (rc) return rc;
        done += nf;
    }
    return MDSF_OK;
}

extern "C" int mdsf_push_density(mdsf_handle* h, const double* d1, int64_t nframes) {
    if (!h || !d1) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    int64_t done = 0;
    double* d_tmp = nullptr;
    CU(cudaMalloc(&d_tmp, sizeof(double) * h->ncell * h->F));
    while (done < nframes) {
        const int nf = (int)std::min<int64_t>(h->F, nframes - done);
        CU(cudaMemcpyAsync(d_tmp, d1 + done * h->ncell, sizeof(double) * h->ncell * nf, cudaMemcpyHostToDevice, h->s_comp));
        const int npairs = (nf + 1) / 2;
        pack_density_kernel<<<grid_for(h->ncell * npairs, 256, h->nsm), 256, 0, h->s_comp>>>(d_tmp, h->d_vol, h->ncell, nf);
        ++h->launches;
        if (h->d_dump) CU(cudaMemcpyAsync(h->d_dump, h->d_vol, sizeof(double2) * h->ncell * npairs, cudaMemcpyDeviceToDevice, h->s_comp));
        int rc = transform_and_accumulate(h, nf, false);
        if (rc) { cudaFree(d_tmp); return rc; }
        h->frames_done += nf;
        h->last_batch_frames = nf;
        done += nf;
    }
    CU(cudaStreamSynchronize(h->s_comp));
    CU(cudaFree(d_tmp));
    return MDSF_OK;
}

extern "C" int mdsf_sync(mdsf_handle* h) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->h_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->s_comp));
    CU(cudaStreamSynchronize(h->s_copy));
    CU(cudaStreamSynchronize(h->s_prep));
    if (h->s_splat) CU(cudaStreamSynchronize(h->s_splat));
    CU(cudaStreamSynchronize(h->s_comp));
    CU(cudaStreamSynchronize(h->s_back));
    if (*h->h_err == 3) {
        *h->h_err = 0;
        cudaMemset(h->d_err, 0, sizeof(int));
        return fail(MDSF_ECUDA, "slab pipeline aborted: a dependency wait exceeded its spin limit");
    }
    if (*h->h_err == 2) {
        *h->h_err = 0;
        cudaMemset(h->d_err, 0, sizeof(int));
        return fail(MDSF_ERANGE, "fixed-point density accumulator overflow (> 2048 peak amplitudes in one cell); use splat mode 1 (owner)");
    }
    if (*h->h_err) {
        *h->h_err = 0;
        cudaMemset(h->d_err, 0, sizeof(int));
        return fail(MDSF_ERANGE, "an atom's Gaussian stamp leaves the padded grid (coordinate more than one box outside the cell, NaN, or half width > Nborder); the reference fails with a numpy shape error here");
    }
    return MDSF_OK;
}

extern "C" int mdsf_export_sf_device(mdsf_handle* h, void* sf_device) {
    if (!h || !sf_device) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const GridParams& gp = h->gp;
    const long long total = (long long)gp.n[0] * gp.n[1] * (gp.n[2] / 2 + 1);
    export_sf_kernel<<<grid_for(total, 256, h->nsm), 256, 0, h->s_comp>>>(h->d_P, (double*)sf_device, h->ax[0].d_rev, h->ax[1].d_rev,
                                                                         h->ax[2].d_rev, gp.n[0], gp.n[1], gp.n[2]);
    ++h->launches;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->s_comp));
    return MDSF_OK;
}

extern "C" int mdsf_read_sf(mdsf_handle* h, double* sf_host) {
    if (!h || !sf_host) return fail(MDSF_EINVAL, "null argument");
    int rc = mdsf_sync(h);
    if (rc) return rc;
    rc = mdsf_export_sf_device(h, h->d_sf);
    if (rc) return rc;
    const GridParams& gp = h->gp;
    const long long total = (long long)gp.n[0] * gp.n[1] * (gp.n[2] / 2 + 1);
    CU(cudaMemcpy(sf_host, h->d_sf, sizeof(double) * total, cudaMemcpyDeviceToHost));
    return MDSF_OK;
}

extern "C" int mdsf_reset(mdsf_handle* h) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(h->d_P, 0, sizeof(double) * h->ncell));
    h->frames_done = 0;
    return MDSF_OK;
}

// ------------------------------------------------------------------------------------------ taps
extern "C" int mdsf_debug_cell_indices(mdsf_handle* h, int64_t frame, int32_t* ir_out) {
    if (!h || !ir_out) return fail(MDSF_EINVAL, "null argument");
    if (frame < 0 || frame >= h->last_batch_frames) return fail(MDSF_EINVAL, "frame %lld not in the last batch (%d frames)", (long long)frame, h->last_batch_frames);
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    std::vector<AtomRec> recs(h->natoms);
    CU(cudaMemcpy(recs.data(), h->d_recs + frame * h->natoms, sizeof(AtomRec) * h->natoms, cudaMemcpyDeviceToHost));
    for (long long a = 0; a < h->natoms; ++a) for (int d = 0; d < 3; ++d) ir_out[a * 3 + d] = recs[a].ir[d];
    return MDSF_OK;
}

extern "C" int mdsf_debug_coords(mdsf_handle* h, int64_t frame, double* r_out) {
    if (!h || !r_out) return fail(MDSF_EINVAL, "null argument");
    if (frame < 0 || frame >= h->last_batch_frames) return fail(MDSF_EINVAL, "frame %lld not in the last batch (%d frames)", (long long)frame, h->last_batch_frames);
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    std::vector<AtomRec> recs(h->natoms);
    CU(cudaMemcpy(recs.data(), h->d_recs + frame * h->natoms, sizeof(AtomRec) * h->natoms, cudaMemcpyDeviceToHost));
    for (long long a = 0; a < h->natoms; ++a) for (int d = 0; d < 3; ++d) r_out[a * 3 + d] = recs[a].r[d];
    return MDSF_OK;
}

extern "C" int mdsf_debug_density(mdsf_handle* h, int64_t frame, double* d1_out) {
    if (!h || !d1_out) return fail(MDSF_EINVAL, "null argument");
    if (!h->d_dump) return fail(MDSF_ESTATE, "create the handle with keep_density=1 to use the density tap");
    if (frame < 0 || frame >= h->last_batch_frames) return fail(MDSF_EINVAL, "frame %lld not in the last batch (%d frames)", (long long)frame, h->last_batch_frames);
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    double* d_tmp = nullptr;
    CU(cudaMalloc(&d_tmp, sizeof(double) * h->ncell));
    unpack_density_kernel<<<grid_for(h->ncell, 256, h->nsm), 256>>>(h->d_dump + (frame / 2) * h->ncell, d_tmp, h->ncell, (int)(frame % 2));
    CU(cudaGetLastError());
    CU(cudaMemcpy(d1_out, d_tmp, sizeof(double) * h->ncell, cudaMemcpyDeviceToHost));
    CU(cudaFree(d_tmp));
    return MDSF_OK;
}

extern "C" int64_t mdsf_kernel_launches(const mdsf_handle* h) { return h ? h->launches : 0; }
extern "C" int64_t mdsf_frames_done(const mdsf_handle* h) { return h ? h->frames_done : 0; }
extern "C" const char* mdsf_fft_path(const mdsf_handle* h) { return (h && h->native_fft) ? "native" : "cufft"; }
extern "C" const char* mdsf_splat_path(const mdsf_handle* h) { return (h && h->scatter) ? "scatter" : ((h && h->tile_atomic) ? "tile" : "owner"); }
extern "C" int mdsf_batch_frames(const mdsf_handle* h) { return h ? h->F : 0; }
extern "C" int mdsf_set_pretransform(mdsf_handle* h, int32_t enabled, double sin_theta, double cos_theta) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    if (enabled && !(sin_theta != 0.0)) return fail(MDSF_EINVAL, "sin(theta) must be non-zero");
    h->mono = enabled ? 1 : 0;
    h->mono_sin = sin_theta;
    h->mono_cos = cos_theta;
    return MDSF_OK;
}
extern "C" int mdsf_pipeline_info(const mdsf_handle* h, int32_t* sms) {
    if (sms) { sms[0] = h ? h->part_sms[0] : 0; sms[1] = h ? h->part_sms[1] : 0; }
    return h ? h->overlap : 0;
}
extern "C" int mdsf_enable_timing(mdsf_handle* h, int32_t on) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    h->timing = on != 0;
    for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
    h->tev.clear();
    h->timed_batches = 0;
    return MDSF_OK;
}
extern "C" int mdsf_stage_ms(mdsf_handle* h, double* out6, int64_t* batches) {
    if (!h || !out6) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < 6; ++i) out6[i] = 0;
    const size_t nb = h->tev.size() / 8;
    for (size_t b = 0; b < nb; ++b) {
        // tv: 0 copy start (s_copy), 1 prep start, 2 binned (prep stream), 6 splat start, 3 splat done (splat stream),
        //     7 y start, 4 y done, 5 end (s_comp).  In overlap mode 3 -> 7 is the wait for the previous batch's x pass.
        cudaEvent_t* tv = &h->tev[b * 8];
        float ms;
        CU(cudaEventElapsedTime(&ms, tv[0], tv[1])); out6[0] += ms;     // copy (overlaps the previous batch's kernels)
        CU(cudaEventElapsedTime(&ms, tv[1], tv[2])); out6[1] += ms;     // prep + bin (overlaps the previous batch when two sets)
        CU(cudaEventElapsedTime(&ms, tv[6], tv[3])); out6[2] += ms;
        CU(cudaEventElapsedTime(&ms, tv[7], tv[4])); out6[3] += ms;
        CU(cudaEventElapsedTime(&ms, tv[4], tv[5])); out6[4] += ms;
        CU(cudaEventElapsedTime(&ms, tv[6], tv[5])); out6[5] += ms;     // compute-stream time of the batch
    }
    if (batches) *batches = (int64_t)nb;
    return MDSF_OK;
}

// Host buffers handed to mdsf_push_frames may be reused once the engine has read them (H2D copy) and, with
// write_back, written them (D2H copy) -- long before the frames are through the pipeline.  A mark is a pair of
// events behind everything queued on the copy / write-back streams so far.
extern "C" int mdsf_input_mark(mdsf_handle* h, int64_t* ticket) {
    if (!h || !ticket) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const int i = (int)(h->next_mark % mdsf_handle::kMarks);
    if (!h->mark_copy[i]) {
        CU(cudaEventCreateWithFlags(&h->mark_copy[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->mark_back[i], cudaEventDisableTiming));
    }
    CU(cudaEventRecord(h->mark_copy[i], h->s_copy));
    CU(cudaEventRecord(h->mark_back[i], h->s_back));
    *ticket = h->next_mark++;
    return MDSF_OK;
}
extern "C" int mdsf_input_wait(mdsf_handle* h, int64_t ticket) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    if (ticket < 0 || ticket >= h->next_mark) return fail(MDSF_EINVAL, "unknown input ticket %lld", (long long)ticket);
    CU(cudaSetDevice(h->device));
    if (h->next_mark - ticket > mdsf_handle::kMarks) return MDSF_OK;     // recycled: a later mark on the same streams was waited for or overwritten
    const int i = (int)(ticket % mdsf_handle::kMarks);
    CU(cudaEventSynchronize(h->mark_copy[i]));
    CU(cudaEventSynchronize(h->mark_back[i]));
    return MDSF_OK;
}

// Device-side stopwatch over everything queued between the two calls: the start event goes on
// the copy stream (every batch begins there), the stop event on the compute stream.
extern "C" int mdsf_timer_start(mdsf_handle* h) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    CU(cudaEventRecord(h->timer0, h->s_copy));
    return MDSF_OK;
}
extern "C" int mdsf_timer_stop(mdsf_handle* h, double* ms_out) {
    if (!h || !ms_out) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->timer1, h->s_comp));
    CU(cudaEventSynchronize(h->timer1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->timer0, h->timer1));
    *ms_out = ms;
    return MDSF_OK;
}
