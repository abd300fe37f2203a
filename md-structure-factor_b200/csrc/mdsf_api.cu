// libmdsf.so -- host side of the C ABI declared in include/mdsf.h.
// One handle = one GPU, four streams (H2D copy, prep+bin, compute, D2H write-back), a double-buffered coordinate
// staging area and two sets of binning buffers, so the copy and the binning of batch b+1 overlap the splat / FFT
// kernels of batch b.
#include <cub/cub.cuh>
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
#include "mdsf_launch.h"
#include "mdsf_prep.cuh"
#include "mdsf_post.cuh"
#include "mdsf_yx.cuh"
#include "mdsf_tma_pass.cuh"

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

static const int kSlots = 2;
static const int kMaxSmem = 227 * 1024 - 1536;      // dynamic shared memory budget: the splat kernels hold up to ~1.1 KB of static shared memory
enum { SPLAT_ORTHO_ = 0, SPLAT_MONO_ = 1, SPLAT_GENERAL_ = 2, SPLAT_DENSITY_ = 3 };     // = the enum of mdsf_splat.cuh

struct AxisPlan {
    FftPlan plan{};
    bool native = false;
    double2* d_tw = nullptr;
    int* d_rev = nullptr;     // frequency index -> position
};

// outputs of the prep/bin stage; two sets so that prep+bin of batch b+1 (stream s_prep) overlap the
// splat/FFT kernels of batch b (stream s_comp)
struct PrepSet {
    AtomRec* recs = nullptr;
    double *tables = nullptr, *tables_alloc = nullptr;
    unsigned *count = nullptr, *start = nullptr;     // [nkeys+1] each: list lengths (K1), list starts -> list ends (K2)
    unsigned *xcnt = nullptr, *perm = nullptr;       // [F*ntx+1] atoms per x bucket -> bucket starts; [F*natoms] walk order of K2
    uint4* prec = nullptr;            // [2 * pair_cap]: 32-byte records {PairRec, PairAux + padding}
    void* cub = nullptr;
    cudaEvent_t ev_binned = nullptr, ev_consumed = nullptr;
    bool used = false;
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
    int splat_mode = 0;               // SPLAT_ORTHO / MONO / GENERAL
    size_t splat_smem = 0;
    double2* d_tws = nullptr;         // per-stage z twiddle tables of the compile-time z path (nullptr: generic stages)
    int tws_n = 0, tws_off = 0;
    // streams / events
    cudaStream_t s_copy = nullptr, s_prep = nullptr, s_comp = nullptr, s_back = nullptr;
    cudaEvent_t ev_h2d[kSlots]{}, ev_free[kSlots]{}, ev_prep[kSlots]{}, ev_back[kSlots]{};
    bool slot_used[kSlots]{};
    int next_slot = 0;
    static const int kMarks = 16;             // mdsf_input_mark / mdsf_input_wait tickets
    cudaEvent_t mark_copy[kMarks] = {}, mark_back[kMarks] = {};
    int64_t next_mark = 0;
    // device buffers
    double *d_amp = nullptr, *d_two = nullptr, *d_ctab = nullptr;
    int *d_halfw = nullptr, *d_ctab_off = nullptr, *d_type = nullptr;
    unsigned* d_toff = nullptr;
    std::vector<int> halfw_host;
    void* d_stage[kSlots]{};
    PrepSet sets[2];
    long long batch_counter = 0;
    AtomRec* d_recs_last = nullptr;   // records of the last batch (parity taps)
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
    long long natoms = 0;
    long long maxpairs_frame = 0;
    int KX = 1, KY = 1;               // most x / y tiles one atom's stamp touches: bin_place_kernel runs KX*KY threads per atom
    unsigned long long pair_cap = 0;
    bool order_atoms = false;         // K2 walks the atoms of a frame by the x tile of their home cell (L2-local list writes)
    bool prep_early = false;          // prep+bin of batch b+1 may start while the splat of batch b still runs
    int mono = 0;                     // K1 applies the monoclinic transform of main_gromacs.py:204-207 first
    double mono_sin = 1.0, mono_cos = 0.0;
    bool tma_y = false, tma_x = false; // TMA-fed persistent pass kernels (512-point axes, lw = 8 layout; mdsf_tma_pass.cuh)
    // fused y -> x pass (mdsf_yx.cuh)
    bool fused_yx = false;
    double2 *d_scratch = nullptr, *d_twsy = nullptr, *d_twsx = nullptr;
    unsigned* d_yxctl = nullptr;
    size_t yxctl_words = 0;
    int yx_grid = 0;
    // y/x pass geometry
    PassGeom pgy{}, pgx{};
    int ntile_y = 1, ntile_x = 1;
    // bookkeeping
    long long launches = 0, frames_done = 0;
    int last_batch_frames = 0;
    bool timing = false;
    std::vector<cudaEvent_t> tev;     // 8 events per timed batch
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
};

// ------------------------------------------------------------------------------------------
static bool factorize(int n, FftPlan& plan, int max_log2, bool allow_big_primes) {
    plan.n = n;
    plan.nstages = 0;
    int e = 0;
    while (n % 2 == 0) { n /= 2; ++e; }
    if (e > 0) {
        const int count = (e + max_log2 - 1) / max_log2, base = e / count, rem = e % count;
        for (int i = 0; i < count; ++i) plan.radix[plan.nstages++] = 1 << (base + (i >= count - rem ? 1 : 0));
    }
    const int odd[] = {3, 5, 7, 11, 13};
    for (int p : odd) {
        if (p > 7 && !allow_big_primes) break;
        while (n % p == 0) {
            if (plan.nstages >= MDSF_MAX_RADIX_STAGES) return false;
            plan.radix[plan.nstages++] = p;
            n /= p;
        }
    }
    return n == 1 && plan.nstages > 0;
}

static int digit_position(int k, int n, const FftPlan& plan, int stage) {
    if (stage >= plan.nstages) return 0;
    const int r = plan.radix[stage], m = n / r;
    return (k % r) * m + digit_position(k / r, m, plan, stage + 1);
}

// z axis: radices <= 8 (the fused splat kernel runs at 64 registers per thread); x / y: up to 16, and 11 / 13
static int build_axis(AxisPlan& ax, int n, bool want_native, bool zaxis) {
    std::vector<int> rev(n);
    ax.native = want_native && factorize(n, ax.plan, zaxis ? 3 : 4, !zaxis);
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

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// per-stage twiddle tables [n2] = w_N^(n2 N/L) of a plan (stages with M = L/R > 1), concatenated in stage order
static int stage_tables(const FftPlan& plan, double2** d_out, int* count) {
    const int N = plan.n;
    std::vector<double2> tws;
    const long double two_pi = 6.283185307179586476925286766559005768L;
    int L = N;
    for (int st = 0; st < plan.nstages; ++st) {
        const int R = plan.radix[st], M = L / R;
        if (M > 1)
            for (int n2 = 0; n2 < M; ++n2) {          // the kernels form the powers w^2 .. w^(R-1) themselves
                const int j = (int)(((long long)n2 * (N / L)) % N);
                const long double a = two_pi * (long double)j / (long double)N;
                tws.push_back(make_double2((double)cosl(a), (double)(-sinl(a))));
            }
        L = M;
    }
    if (tws.empty()) tws.push_back(make_double2(1.0, 0.0));
    if (count) *count = (int)tws.size();
    CU(cudaMalloc(d_out, sizeof(double2) * tws.size()));
    CU(cudaMemcpy(*d_out, tws.data(), sizeof(double2) * tws.size(), cudaMemcpyHostToDevice));
    return MDSF_OK;
}

// Slab geometry of the splat for `sub` lists per warp (1: one slab of 256 >> lcol cells per warp, 2: two half-width
// slabs side by side), its shared memory, and whether the z twiddle tables get their own region (prefetched while the
// splat runs) -- they do when two CTAs still fit an SM.
static void configure_splat(mdsf_handle* h, int sub) {
    GridParams& gp = h->gp;
    // z-lane kernel (separable ucell): one full-warp slab per list, lanes walk z (mdsf_splat.cuh); MDSF_ZLANE=0 keeps the
    // (row, z lane) x TX-column kernel with `sub` half-warp lists per warp
    // (measured, splat per step: c1 8.24 -> 4.97 ms, c3 19.7 -> 19.7 ms, c4 43.4 -> 42.5 ms; 4x4-column tiles lose:
    // c2 7.19 -> 7.98 ms, so they keep the half-warp lists)
    gp.zlane = (gp.separable && env_int("MDSF_ZLANE", gp.lcol == 4 ? 0 : 1) != 0) ? 1 : 0;
    if (gp.zlane) sub = 1;
    gp.sub = sub;
    gp.zw = (256 >> gp.lcol) / sub;
    gp.nslab = (gp.n[2] + gp.zw - 1) / gp.zw;
    h->splat_smem = mdsf_splat_smem(gp.lcol, sub, gp.nzp, gp.n[2], gp.zlane);
    h->tws_off = 0;
    if (h->d_tws) {
        const size_t base = (h->splat_smem + 15) / 16 * 16, need = base + sizeof(double2) * (size_t)h->tws_n;
        if (2 * (need + 1024) <= (size_t)kMaxSmem + 1024) { h->tws_off = (int)base; h->splat_smem = need; }
    }
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
        if (cfg->n[d] > (1 << 20)) return fail(MDSF_EINVAL, "grid size n[%d]=%d too large", d, cfg->n[d]);
    }
    if (cfg->ntypes < 1 || !cfg->amp || !cfg->two_sig2 || !cfg->halfw) return fail(MDSF_EINVAL, "type tables missing");
    if (cfg->nborder < 0) return fail(MDSF_EINVAL, "negative Nborder");
    for (int t = 0; t < cfg->ntypes * 3; ++t) {
        if (2 * cfg->halfw[t] > MDSF_MAX_STAMP) return fail(MDSF_EINVAL, "stamp of %d cells exceeds %d", 2 * cfg->halfw[t], MDSF_MAX_STAMP);
        if (cfg->halfw[t] < 0 || cfg->halfw[t] > cfg->nborder) return fail(MDSF_EINVAL, "half width %d outside [0, Nborder=%d]", cfg->halfw[t], cfg->nborder);
    }
    if (cfg->coord_dtype != MDSF_F32 && cfg->coord_dtype != MDSF_F64) return fail(MDSF_EINVAL, "bad coord_dtype");
    if (cfg->arith_dtype != MDSF_F32 && cfg->arith_dtype != MDSF_F64) return fail(MDSF_EINVAL, "bad arith_dtype");
    if (cfg->coord_dtype == MDSF_F64 && cfg->arith_dtype == MDSF_F32) return fail(MDSF_EINVAL, "float64 coordinates never promote to float32");
    if (cfg->splat_mode != 0) return fail(MDSF_EINVAL, "splat_mode must be 0: the engine has one splat (register-tiled fixed-point owner-computes)");

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
    gp.fold_mode = cfg->fold_mode;
    // z decouples when ucell[2][0]=ucell[2][1]=0 (b_z feeds only c_2) and ucell[0][2]=ucell[1][2]=0
    gp.separable = (gp.u[6] == 0.0 && gp.u[7] == 0.0 && gp.u[2] == 0.0 && gp.u[5] == 0.0) ? 1 : 0;
    gp.cxx = gp.u[0] * gp.u[0] + gp.u[1] * gp.u[1];
    gp.cyy = gp.u[3] * gp.u[3] + gp.u[4] * gp.u[4];
    gp.gxy = gp.u[0] * gp.u[3] + gp.u[1] * gp.u[4];
    gp.czz = gp.u[8] * gp.u[8];
    h->ncell = (long long)gp.n[0] * gp.n[1] * gp.n[2];
    h->prep_early = env_int("MDSF_PREP_EARLY", 0) != 0;
    h->splat_mode = !gp.separable ? SPLAT_GENERAL_ : (gp.gxy != 0.0 ? SPLAT_MONO_ : SPLAT_ORTHO_);

    // ---- FFT plans.  Grids the fused y -> x kernel covers use radices <= 8 on every axis (its stage lists), a z-chunked
    // volume layout of width 4 and the L2-resident hand-over; everything else runs separate y and x passes.
    const bool want_native = cfg->fft_mode != MDSF_FFT_CUFFT;
    // the fused kernel (opt-in, MDSF_FUSED_YX=1): a third of the HBM traffic of separate passes, but its per-tile
    // hand-shakes leave it latency-bound (c3: 15.1 ms per 16 frames against 13.4 ms for the register-staged passes and
    // 10 ms for the TMA-fed ones)
    h->fused_yx = want_native && mdsf_yx_supported(gp.n[1], gp.n[0]) && gp.n[2] % MDSF_YX_W == 0 && gp.n[2] <= 2048 &&
                  env_int("MDSF_FUSED_YX", 0) != 0 && env_int("MDSF_LAYOUT_W", 0) == 0;
    // TMA-fed persistent passes: 512-point y and / or x axis, z-chunked layout of width 8
    // (MDSF_TMA_PASS: bit 0 = y pass, bit 1 = x pass)
    const int tma_want = env_int("MDSF_TMA_PASS", 3);
    const bool tma_ok = want_native && !h->fused_yx && gp.n[2] % 8 == 0 &&
                        (env_int("MDSF_LAYOUT_W", 0) == 0 || env_int("MDSF_LAYOUT_W", 0) == 8);
    h->tma_y = tma_ok && (tma_want & 1) && gp.n[1] == MDSF_TP_N;
    h->tma_x = tma_ok && (tma_want & 2) && gp.n[0] == MDSF_TP_N;
    for (int d = 0; d < 3; ++d) {
        int rc = build_axis(h->ax[d], gp.n[d], want_native, d == 2 || h->fused_yx);
        if (rc) return rc;
    }
    h->native_fft = h->ax[0].native && h->ax[1].native && h->ax[2].native;
    if (gp.n[2] > 2048 || gp.n[1] > 2048 || gp.n[0] > 2048) h->native_fft = false;
    if (!h->native_fft) h->fused_yx = h->tma_y = h->tma_x = false;
    {   // the TMA-fed kernels hard-wire the stage list 8 * 8 * 8 (and its digit-reversed output order)
        auto is888 = [](const FftPlan& p) { return p.nstages == 3 && p.radix[0] == 8 && p.radix[1] == 8 && p.radix[2] == 8; };
        if (!is888(h->ax[1].plan)) h->tma_y = false;
        if (!is888(h->ax[0].plan)) h->tma_x = false;
    }
    if (!h->native_fft) {
        if (cfg->fft_mode == MDSF_FFT_NATIVE)
            return fail(MDSF_EINVAL, "grid %dx%dx%d has a prime factor > 13 (> 7 in z) or an axis longer than 2048; native FFT unavailable", gp.n[0], gp.n[1], gp.n[2]);
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
        if (last == 8) gp.pad_shift = 3; else if (last == 4) gp.pad_shift = 2;
    }
    gp.nzp = (gp.n[2] + (gp.pad_shift < 31 ? (gp.n[2] >> gp.pad_shift) : 0)) | 1;   // odd: columns start in different banks

    // ---- splat tile: 2^lcol columns over all z, 16 B per cell in shared memory; two CTAs per SM when the tile allows
    if (cfg->tile_x > 0 && cfg->tile_y > 0) {
        int l = -1;
        for (int k = 2; k <= 5; ++k) if (cfg->tile_x == (1 << ((k + 1) / 2)) && cfg->tile_y == (1 << (k / 2))) l = k;
        if (l < 0) return fail(MDSF_EINVAL, "tile %dx%d: supported tiles are 2x2, 4x2, 4x4 and 8x4 columns", cfg->tile_x, cfg->tile_y);
        gp.lcol = l;
    } else {
        gp.lcol = 5;
        const size_t tile_budget = (size_t)env_int("MDSF_TILE_KB", 112) * 1024;      // 112 KB: two CTAs per SM
        while (gp.lcol > 2 && mdsf_splat_smem(gp.lcol, 1, gp.nzp, gp.n[2]) > tile_budget) --gp.lcol;
    }
    gp.zilv = 0;
    if (h->native_fft && mdsf_zspec_length(gp.n[2]) && env_int("MDSF_ZILV", 1) != 0) {
        // compile-time z stages on the interleaved tile: [col][nz + 1] cells of 16 bytes (odd column stride, no padding inside a column)
        gp.zilv = 1;
        gp.pad_shift = 31;
        gp.nzp = gp.n[2] + 1;
    }
    if (mdsf_splat_smem(gp.lcol, 1, gp.nzp, gp.n[2]) > (size_t)kMaxSmem)
        return fail(MDSF_EINVAL, "grid too long in z (%d) for the column-tile splat", gp.n[2]);
    {
        const int TX = 1 << ((gp.lcol + 1) / 2), TY = 1 << (gp.lcol / 2);
        gp.ntx = (gp.n[0] + TX - 1) / TX;
        gp.nty = (gp.n[1] + TY - 1) / TY;
    }
    // ---- compile-time z stages: per-stage twiddle tables [n2] = w^(n2 N/L)
    if (gp.zilv) {
        int rc = stage_tables(h->ax[2].plan, &h->d_tws, &h->tws_n);
        if (rc) return rc;
    }
    configure_splat(h, 1);

    // ---- volume layout: plain [x][y][z], or z-chunked [z/lw][x][y][lw] (MDSF_LAYOUT_W = 4 / 8; native FFT only)
    gp.lw = h->fused_yx ? MDSF_YX_W : ((h->tma_y || h->tma_x) ? 8 : gp.n[2]);
    // (large planes: x rows of the plain layout are Ny*Nz*16 bytes apart -- 4 MB at 512^3, one TLB entry per row; the
    // chunked layout of width 8 keeps 128-byte rows and cuts that stride to Ny*128 bytes: c3 x pass 5.62 -> 5.35 ms, c4 +4 %)
    if (!h->fused_yx && h->native_fft && gp.n[2] % 8 == 0 && (long long)gp.n[0] * gp.n[1] >= 512LL * 512) gp.lw = 8;
    {
        const int want = env_int("MDSF_LAYOUT_W", 0);
        if (want > 0 && h->native_fft) {
            if ((want != 4 && want != 8) || gp.n[2] % want) return fail(MDSF_EINVAL, "MDSF_LAYOUT_W=%d: must be 4 or 8 and divide Nz=%d", want, gp.n[2]);
            gp.lw = want;
        }
    }
    gp.nch = gp.n[2] / gp.lw;

    // ---- batch size
    int F = cfg->batch_frames;
    if (F <= 0) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const double per_pair = (double)h->ncell * 16.0 * (cfg->keep_density ? 2 : 1);
        const double budget = std::min(18.0e9, (double)free_b * 0.25);
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
        // fixed-point accumulators of the splat: one term is < 2^52 LSBs (amax < 2^e), 2^11 of headroom per cell
        double amax = 0;
        for (int t = 0; t < nt; ++t) amax = std::max(amax, cfg->amp[t]);
        int ex = 0;
        (void)std::frexp(amax, &ex);                       // amax < 2^ex
        gp.fx_scale = std::ldexp(1.0, 52 - ex);
        gp.fx_inv = std::ldexp(1.0, ex - 52);
    }
    h->halfw_host.assign(cfg->halfw, cfg->halfw + nt * 3);
    h->tt.amp = h->d_amp; h->tt.two_sig2 = h->d_two; h->tt.halfw = h->d_halfw;
    h->tt.ctab = nullptr; h->tt.ctab_off = nullptr; h->tt.toff = nullptr;
    if (h->splat_mode == SPLAT_MONO_) {
        // cross-term table of every type: C[i][j] = exp(-2 gxy dx dy i j / (2 sigma^2))
        // columns outside a stamp index the tables out of a type's range (their EX / EY factor is an exact zero, so the
        // value only has to be finite): pads of ones in front and behind cover the largest excursion, 8 rows + 8 entries
        std::vector<int> off(nt);
        int amax2 = 2;
        for (int t = 0; t < nt; ++t) amax2 = std::max(amax2, 2 * cfg->halfw[t * 3 + 1]);
        const size_t cpad = (size_t)8 * amax2 + 8;
        std::vector<double> ctab(cpad, 1.0);
        for (int t = 0; t < nt; ++t) {
            off[t] = (int)ctab.size();
            const int ax2 = 2 * cfg->halfw[t * 3], ay2 = 2 * cfg->halfw[t * 3 + 1];
            for (int i = 0; i < ax2; ++i)
                for (int j = 0; j < ay2; ++j)
                    ctab.push_back(std::exp(-(2.0 * gp.gxy * gp.dr[0] * gp.dr[1] * (double)i * (double)j) / cfg->two_sig2[t]));
        }
        ctab.insert(ctab.end(), cpad, 1.0);
        if (ctab.size() >= (1u << 30)) return fail(MDSF_EINVAL, "cross-term tables too large");
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
    if (cfg->keep_density) CU(cudaMalloc(&h->d_dump, sizeof(double2) * h->ncell * npairs));
    CU(cudaMalloc(&h->d_P, sizeof(double) * h->ncell));
    CU(cudaMemset(h->d_P, 0, sizeof(double) * h->ncell));
    CU(cudaMalloc(&h->d_sf, sizeof(double) * (long long)gp.n[0] * gp.n[1] * (gp.n[2] / 2 + 1)));
    CU(cudaMalloc(&h->d_err, sizeof(int)));
    CU(cudaMemset(h->d_err, 0, sizeof(int)));
    CU(cudaMallocHost(&h->h_err, sizeof(int)));
    *h->h_err = 0;

    CU(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_back, cudaStreamNonBlocking));
    {
        // prep+bin of batch b+1 run underneath the HBM-bound y/x passes of batch b; at equal priority their small
        // kernels queue behind the passes' CTAs and finish after them (a gap before the next splat)
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&h->s_prep, cudaStreamNonBlocking, hi));
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
        // y / x pass tiles: [n][W] complex in shared memory; W = 8 keeps 128-byte rows, narrower only when the
        // tile would not fit (axes > 1024 points); the chunked layout fixes W = lw
        auto pick = [&](int n, int arrays) {
            int W = 8;
            while (W > 1 && (size_t)n * W * 8 * arrays + (size_t)n * 16 > (size_t)kMaxSmem - 2048) W >>= 1;
            return W;
        };
        const bool chunked = gp.lw != gp.n[2];
        const int Wy = chunked ? gp.lw : pick(gp.n[1], 2), Wx = chunked ? gp.lw : pick(gp.n[0], 3);
        h->pgy.ncell = h->pgx.ncell = h->ncell;
        h->pgy.nz = h->pgx.nz = gp.n[2];
        h->pgy.W = Wy; h->pgx.W = Wx;
        if (chunked) {
            h->pgy.os = (long long)gp.n[1] * gp.lw; h->pgy.rs = gp.lw; h->pgy.cs = (long long)gp.n[0] * gp.n[1] * gp.lw;
            h->pgx.os = gp.lw; h->pgx.rs = (long long)gp.n[1] * gp.lw; h->pgx.cs = h->pgy.cs;
        } else {
            h->pgy.os = (long long)gp.n[1] * gp.n[2]; h->pgy.rs = gp.n[2]; h->pgy.cs = Wy;
            h->pgx.os = gp.n[2]; h->pgx.rs = (long long)gp.n[1] * gp.n[2]; h->pgx.cs = Wx;
        }
        h->ntile_y = (gp.n[2] + Wy - 1) / Wy;
        h->ntile_x = (gp.n[2] + Wx - 1) / Wx;
        CU(mdsf_pass_configure());
    }
    CU(mdsf_splat_configure());
    if (h->tma_y || h->tma_x) {
        if (gp.lw != 8) h->tma_y = h->tma_x = false;
        else CU(mdsf_tma_pass_configure());
    }
    if (h->fused_yx || h->tma_y || h->tma_x) {
        int rc = stage_tables(h->ax[1].plan, &h->d_twsy, nullptr);
        if (rc) return rc;
        rc = stage_tables(h->ax[0].plan, &h->d_twsx, nullptr);
        if (rc) return rc;
    }
    if (h->fused_yx) {
        const long long slab = (long long)gp.n[0] * gp.n[1] * MDSF_YX_W;
        CU(cudaMalloc(&h->d_scratch, sizeof(double2) * slab * MDSF_YX_RING));
        h->yxctl_words = 1 + (size_t)2 * gp.nch * npairs + (size_t)gp.nch * gp.n[1];
        CU(cudaMalloc(&h->d_yxctl, sizeof(unsigned) * h->yxctl_words));
        const int per_sm = mdsf_yx_blocks_per_sm(gp.n[1], gp.n[0]);
        if (per_sm < 1) return fail(MDSF_ECUDA, "fused y/x kernel does not fit an SM");
        h->yx_grid = per_sm * h->nsm;
    }
    *out = h;
    return MDSF_OK;
}

extern "C" int mdsf_destroy(mdsf_handle* h) {
    if (!h) return MDSF_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->cufft_plan) cufftDestroy(h->cufft_plan);
    void* bufs[] = {h->d_scratch, h->d_twsy, h->d_twsx, h->d_yxctl, h->d_tws, h->d_ctab, h->d_ctab_off, h->d_toff, h->d_amp, h->d_two, h->d_halfw, h->d_type, h->d_stage[0], h->d_stage[1],
                    h->d_vol, h->d_dump, h->d_P, h->d_sf, h->d_err};
    for (void* b : bufs) if (b) cudaFree(b);
    for (int d = 0; d < 3; ++d) { if (h->ax[d].d_tw) cudaFree(h->ax[d].d_tw); if (h->ax[d].d_rev) cudaFree(h->ax[d].d_rev); }
    for (auto& ps : h->sets) {
        void* pb[] = {ps.recs, ps.tables_alloc, ps.count, ps.start, ps.xcnt, ps.perm, ps.prec, ps.cub};
        for (void* b : pb) if (b) cudaFree(b);
        if (ps.ev_binned) cudaEventDestroy(ps.ev_binned);
        if (ps.ev_consumed) cudaEventDestroy(ps.ev_consumed);
    }
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
    if (h->s_prep) cudaStreamDestroy(h->s_prep);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    if (h->s_back) cudaStreamDestroy(h->s_back);
    delete h;
    return MDSF_OK;
}

extern "C" int mdsf_set_atoms(mdsf_handle* h, int64_t natoms, const int32_t* type_id) {
    if (!h || !type_id) return fail(MDSF_EINVAL, "null argument");
    if (natoms < 1 || natoms >= MDSF_MAX_ATOMS) return fail(MDSF_EINVAL, "natoms=%lld outside [1, %d)", (long long)natoms, MDSF_MAX_ATOMS);
    if (h->d_type) return fail(MDSF_ESTATE, "atoms already set on this handle");
    CU(cudaSetDevice(h->device));
    const int nt = h->cfg.ntypes;
    {
        // two lists per warp when the stamps are short against a full-warp slab (most of its cells would add zeros)
        double zsum = 0;
        for (int64_t a = 0; a < natoms; ++a) {
            if (type_id[a] < 0 || type_id[a] >= nt) return fail(MDSF_EINVAL, "type_id[%lld]=%d out of range", (long long)a, type_id[a]);
            zsum += 2.0 * h->halfw_host[type_id[a] * 3 + 2];
        }
        int sub = (zsum / (double)natoms <= 0.5 * (256 >> h->gp.lcol)) ? 2 : 1;
        sub = env_int("MDSF_SUB", sub);
        if (sub != 1 && sub != 2) return fail(MDSF_EINVAL, "MDSF_SUB must be 1 or 2");
        configure_splat(h, sub);
    }
    const GridParams& g0 = h->gp;
    const int TX = 1 << ((g0.lcol + 1) / 2), TY = 1 << (g0.lcol / 2);
    // worst-case (image, tile, slab) pairs per atom of each type: exact maximum over every admissible cell index,
    // per dimension (z: both fold shifts of a padding segment, -+Nz and the corner rule's +-Nborder)
    std::vector<long long> bound(nt);
    h->KX = h->KY = 1;
    for (int t = 0; t < nt; ++t) {
        long long m[3] = {1, 1, 1};
        for (int d = 0; d < 3; ++d) {
            const int A = h->halfw_host[t * 3 + d], N = g0.n[d], tl = d == 0 ? TX : (d == 1 ? TY : g0.zw);
            for (int ir = A - g0.nb; ir <= N + g0.nb - A; ++ir) {
                m[d] = std::max<long long>(m[d], stamp_bins_1d(ir, A, N, tl, N, -N));
                if (d == 2 && g0.fold_mode == 0) {
                    m[d] = std::max<long long>(m[d], stamp_bins_1d(ir, A, N, tl, g0.nb, -N));
                    m[d] = std::max<long long>(m[d], stamp_bins_1d(ir, A, N, tl, N, -g0.nb));
                }
            }
        }
        bound[t] = m[0] * m[1] * m[2];
        h->KX = std::max(h->KX, (int)m[0]);
        h->KY = std::max(h->KY, (int)m[1]);
    }
    long long maxpairs = 0;
    std::vector<unsigned> toff(natoms);
    long long tstride = 0;
    for (int64_t a = 0; a < natoms; ++a) {
        if (type_id[a] < 0 || type_id[a] >= nt) return fail(MDSF_EINVAL, "type_id[%lld]=%d out of range", (long long)a, type_id[a]);
        maxpairs += bound[type_id[a]];
        toff[a] = (unsigned)tstride;
        const int* hw = &h->halfw_host[type_id[a] * 3];
        tstride += table_doubles(g0.lcol, hw[0], hw[1], hw[2]);
    }
    // pair lists and factor tables scale with the batch: shrink it until both fit 32-bit offsets and ~24 GB
    while (h->F > 2 && (tstride * h->F >= (1LL << 31) - 64 || maxpairs * h->F >= (1LL << 32) - 2 ||
                        (double)maxpairs * h->F * 32.0 > 12.0e9) && h->cfg.batch_frames <= 0)
        h->F -= 2;
    if (tstride * h->F >= (1LL << 31) - 64) return fail(MDSF_EINVAL, "factor tables overflow 31-bit offsets; lower batch_frames");
    h->gp.tstride = tstride;
    h->natoms = natoms;
    h->gp.natoms = (int)natoms;
    h->maxpairs_frame = maxpairs;
    const long long cap = maxpairs * h->F;
    if (cap >= (1LL << 32) - 2) return fail(MDSF_EINVAL, "pair capacity %lld overflows 32-bit offsets; lower batch_frames", cap);
    h->pair_cap = (unsigned long long)std::max(1LL, cap);
    const long long nkeys = (long long)h->F * g0.ntx * g0.nty * g0.nslab;
    if (nkeys >= (1LL << 31)) return fail(MDSF_EINVAL, "too many (tile, slab) lists per batch; lower batch_frames");

    // K2 walks the atoms in x-bucket order (order_atoms_kernel) when the lists of one frame are too large to sit in L2
    // while file-order atoms fill them (c3: ~220 MB per frame; c2's 27 MB do not need it)
    h->order_atoms = env_int("MDSF_ORDER_ATOMS", (double)maxpairs * 32.0 > 96.0e6 ? 1 : 0) != 0 && (long long)natoms * h->F < (1LL << 32);
    CU(cudaMalloc(&h->d_type, sizeof(int) * natoms));
    CU(cudaMemcpy(h->d_type, type_id, sizeof(int) * natoms, cudaMemcpyHostToDevice));
    for (int s = 0; s < kSlots; ++s) CU(cudaMalloc(&h->d_stage[s], h->csize * 3 * natoms * h->F));
    CU(cudaMalloc(&h->d_toff, sizeof(unsigned) * natoms));
    CU(cudaMemcpy(h->d_toff, toff.data(), sizeof(unsigned) * natoms, cudaMemcpyHostToDevice));
    h->tt.toff = h->d_toff;
    {
        size_t b = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b, (unsigned*)nullptr, (unsigned*)nullptr, (long long)nkeys + 1, h->s_prep);
        h->cub_bytes = b + 256;
    }
    for (int p = 0; p < 2; ++p) {
        PrepSet& ps = h->sets[p];
        CU(cudaMalloc(&ps.recs, sizeof(AtomRec) * natoms * h->F));
        // 64 doubles of slack on both sides: zero-filled cp.async copies of EZ entries outside a record's window still
        // carry an address up to one slab width before / behind the atom's block
        CU(cudaMalloc(&ps.tables_alloc, sizeof(double) * (std::max(1LL, tstride * h->F) + 128)));
        CU(cudaMemset(ps.tables_alloc, 0, sizeof(double) * (std::max(1LL, tstride * h->F) + 128)));
        ps.tables = ps.tables_alloc + 64;
        CU(cudaMalloc(&ps.count, sizeof(unsigned) * (nkeys + 2)));
        CU(cudaMalloc(&ps.start, sizeof(unsigned) * (nkeys + 2)));
        CU(cudaMalloc(&ps.prec, 2 * sizeof(uint4) * h->pair_cap));
        if (h->order_atoms) {
            CU(cudaMalloc(&ps.xcnt, sizeof(unsigned) * ((size_t)h->F * g0.ntx + 2)));
            CU(cudaMalloc(&ps.perm, sizeof(unsigned) * (size_t)natoms * h->F));
        }
        CU(cudaMalloc(&ps.cub, h->cub_bytes));
        CU(cudaEventCreateWithFlags(&ps.ev_binned, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ps.ev_consumed, cudaEventDisableTiming));
    }
    h->d_recs_last = h->sets[0].recs;
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
// y / x FFT passes + |C|^2 accumulation of the pair volumes in d_vol (nf frames -> (nf+1)/2 pairs); the z pass ran
// inside the splat kernel.  Library path: the volumes hold densities, cuFFT does all three axes.
static int transform_and_accumulate(mdsf_handle* h, int nf, cudaEvent_t* tv = nullptr) {
    const GridParams& gp = h->gp;
    const int npairs = (nf + 1) / 2;
    if (tv) CU(cudaEventRecord(tv[3], h->s_comp));
    if (h->native_fft && h->fused_yx) {
        CU(cudaMemsetAsync(h->d_yxctl, 0, sizeof(unsigned) * h->yxctl_words, h->s_comp));
        YXParams yp{};
        yp.vol = h->d_vol; yp.scratch = h->d_scratch; yp.P = h->d_P; yp.twy = h->d_twsy; yp.twx = h->d_twsx;
        yp.nch = gp.nch; yp.npairs = npairs; yp.ctl = h->d_yxctl; yp.err = h->d_err;
        CU(mdsf_launch_yx(gp.n[1], gp.n[0], yp, h->yx_grid, h->s_comp));
        ++h->launches;
        if (tv) CU(cudaEventRecord(tv[4], h->s_comp));
    } else if (h->native_fft) {
        cudaError_t ce = cudaSuccess;
        PassArgs a{};
        a.vol = h->d_vol; a.P = h->d_P; a.npairs = npairs; a.scratch = nullptr;
        TPParams tp{};
        tp.vol = h->d_vol; tp.P = h->d_P; tp.nx = gp.n[0]; tp.ny = gp.n[1]; tp.nch = gp.nch; tp.npairs = npairs;
        int nl = 1;
        if (h->tma_y) {
            tp.tws = h->d_twsy;
            ce = mdsf_launch_tma_pass(false, tp, h->nsm, h->s_comp);
            if (ce != cudaSuccess) nl = -1;
        } else {
            a.plan = &h->ax[1].plan; a.tw = h->ax[1].d_tw; a.pg = h->pgy; a.nouter = gp.n[0]; a.ntile = h->ntile_y;
            nl = mdsf_launch_pass_y(a, h->s_comp, &ce);
        }
        if (nl < 0) return fail(MDSF_ECUDA, "y pass launch failed: %s", cudaGetErrorString(ce));
        h->launches += nl;
        if (tv) CU(cudaEventRecord(tv[4], h->s_comp));
        nl = 1;
        if (h->tma_x) {
            tp.tws = h->d_twsx;
            ce = mdsf_launch_tma_pass(true, tp, h->nsm, h->s_comp);
            if (ce != cudaSuccess) nl = -1;
        } else {
            a.plan = &h->ax[0].plan; a.tw = h->ax[0].d_tw; a.pg = h->pgx; a.nouter = gp.n[1]; a.ntile = h->ntile_x;
            nl = mdsf_launch_pass_x(a, h->s_comp, &ce);
        }
        if (nl < 0) return fail(MDSF_ECUDA, "x pass launch failed: %s", cudaGetErrorString(ce));
        h->launches += nl;
    } else {
        if (npairs != h->cufft_batch)    // partial last batch: zero the unused pair volumes and transform the whole batch
            CU(cudaMemsetAsync(h->d_vol + (long long)npairs * h->ncell, 0, sizeof(double2) * h->ncell * (h->cufft_batch - npairs), h->s_comp));
        CF(cufftExecZ2Z(h->cufft_plan, (cufftDoubleComplex*)h->d_vol, (cufftDoubleComplex*)h->d_vol, CUFFT_FORWARD));
        if (tv) CU(cudaEventRecord(tv[4], h->s_comp));
        accumulate_power_kernel<<<grid_for(h->ncell, 256, h->nsm), 256, 0, h->s_comp>>>(h->d_vol, h->d_P, h->ncell, npairs);
        ++h->launches;
    }
    CU(cudaGetLastError());
    return MDSF_OK;
}

template <typename C, typename P>
static void launch_prep(mdsf_handle* h, cudaStream_t st, void* stage, PrepSet& ps, const BatchScales& sc, int nf, long long wlo, long long whi) {
    const long long total = (long long)nf * h->natoms;
    prep_atoms_kernel<C, P><<<grid_for(total, 256, h->nsm), 256, 0, st>>>(
        (C*)stage, h->d_type, ps.recs, ps.tables, h->gp, h->tt, sc, nf, wlo, whi, h->d_err, ps.xcnt,
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
    const int p = (int)(h->batch_counter++ & 1);
    PrepSet& ps = h->sets[p];
    PrepSet& other = h->sets[1 - p];
    cudaStream_t sp = h->s_prep;
    cudaEvent_t* tv = nullptr;
    if (h->timing) {
        for (int i = 0; i < 8; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); h->tev.push_back(e); }
        tv = &h->tev[h->tev.size() - 8];
        CU(cudaEventRecord(tv[0], h->s_copy));
    }
    CU(cudaMemcpyAsync(h->d_stage[slot], src, bytes, cudaMemcpyDefault, h->s_copy));
    CU(cudaEventRecord(h->ev_h2d[slot], h->s_copy));
    CU(cudaStreamWaitEvent(sp, h->ev_h2d[slot], 0));
    if (ps.used) CU(cudaStreamWaitEvent(sp, ps.ev_consumed, 0));      // the splat of batch b-2 has read this set
    // start after the splat of batch b-1: prep+bin then overlap its HBM-bound y/x passes instead of fighting the
    // issue-bound splat kernel for the SMs
    if (other.used && !h->prep_early) CU(cudaStreamWaitEvent(sp, other.ev_consumed, 0));
    if (tv) CU(cudaEventRecord(tv[1], sp));

    BatchScales sc;
    for (int f = 0; f < nf; ++f) for (int d = 0; d < 3; ++d) sc.a[f][d] = scale[f * 3 + d];
    // an odd batch gets a phantom last frame with empty lists (imaginary part of the last pair)
    const unsigned nkeys = (unsigned)((nf + (nf & 1)) * gp.ntx * gp.nty * gp.nslab);
    CU(cudaMemsetAsync(ps.count, 0, sizeof(unsigned) * (nkeys + 1), sp));
    if (ps.xcnt) CU(cudaMemsetAsync(ps.xcnt, 0, sizeof(unsigned) * ((size_t)nf * gp.ntx + 1), sp));
    const bool c32 = h->cfg.coord_dtype == MDSF_F32, p32 = h->cfg.arith_dtype == MDSF_F32;
    if (c32 && p32) launch_prep<float, float>(h, sp, h->d_stage[slot], ps, sc, nf, wlo, whi);
    else if (c32) launch_prep<float, double>(h, sp, h->d_stage[slot], ps, sc, nf, wlo, whi);
    else launch_prep<double, double>(h, sp, h->d_stage[slot], ps, sc, nf, wlo, whi);
    ++h->launches;
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev_prep[slot], sp));
    CU(cudaEventRecord(h->ev_free[slot], sp));          // the staging slot is only read (and rewritten) by K1
    if (write_back) {
        CU(cudaStreamWaitEvent(h->s_back, h->ev_prep[slot], 0));
        CU(cudaMemcpyAsync(src, h->d_stage[slot], bytes, cudaMemcpyDefault, h->s_back));
    }
    CU(cudaEventRecord(h->ev_back[slot], h->s_back));

    // K2: list starts = exclusive scan of the K1 counts (in place); placement claims slots from that same array, which
    // leaves the list ENDS in it: list k = [k ? end[k-1] : 0, end[k])
    size_t cb = h->cub_bytes;
    const long long total = (long long)nf * h->natoms;
    if (ps.xcnt) {      // K2a: bucket starts (in place), then the walk order
        cub::DeviceScan::ExclusiveSum(ps.cub, cb, ps.xcnt, ps.xcnt, (long long)nf * gp.ntx + 1, sp);
        order_atoms_kernel<<<grid_for(total, 256, h->nsm), 256, 0, sp>>>(ps.recs, ps.xcnt, ps.perm, gp, nf);
        h->launches += 2;
    }
    bin_count_kernel<<<grid_for(total, 256, h->nsm), 256, 0, sp>>>(ps.recs, ps.perm, ps.count, gp, h->tt, nf);
    ++h->launches;
    cub::DeviceScan::ExclusiveSum(ps.cub, cb, ps.count, ps.start, (long long)nkeys + 1, sp);
    bin_place_kernel<<<grid_for(total, 256, h->nsm), 256, 0, sp>>>(ps.recs, ps.perm, ps.start, ps.prec, gp, h->tt, nf, nkeys, h->pair_cap, h->d_err);
    h->launches += 2;
    CU(cudaGetLastError());
    if (tv) CU(cudaEventRecord(tv[2], sp));
    CU(cudaEventRecord(ps.ev_binned, sp));
    CU(cudaStreamWaitEvent(h->s_comp, ps.ev_binned, 0));
    if (tv) CU(cudaEventRecord(tv[6], h->s_comp));

    // K3: splat (+ fused z FFT on the native path)
    const int npairs = (nf + 1) / 2;
    SplatArgs sa{};
    sa.prec = ps.prec; sa.start = ps.start; sa.recs = ps.recs; sa.tables = ps.tables;
    sa.src_density = nullptr; sa.nframes = nf; sa.vol = h->d_vol; sa.dens_dump = h->d_dump; sa.gp = gp; sa.tt = h->tt;
    sa.zplan = h->ax[2].plan; sa.twz = h->ax[2].d_tw; sa.err_flag = h->d_err;
    sa.tws = h->d_tws; sa.tws_n = h->tws_n; sa.tws_off = h->tws_off;
    CU(mdsf_launch_splat(gp.lcol, h->splat_mode, h->native_fft, dim3(gp.ntx * gp.nty, npairs), h->splat_smem, h->s_comp, sa));
    ++h->launches;
    CU(cudaEventRecord(ps.ev_consumed, h->s_comp));
    ps.used = true;
    h->d_recs_last = ps.recs;
    int rc = transform_and_accumulate(h, nf, tv);
    if (rc) return rc;
    if (tv) CU(cudaEventRecord(tv[5], h->s_comp));
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

// RANDOM_NOISE mode (dens.py:279-280): ready-made densities go through the same tile kernel (its list walk replaced
// by a load of the density columns), i.e. through the same fused z pass, then the y / x passes.
extern "C" int mdsf_push_density(mdsf_handle* h, const double* d1, int64_t nframes) {
    if (!h || !d1) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const GridParams& gp = h->gp;
    int64_t done = 0;
    double* d_tmp = nullptr;
    CU(cudaMalloc(&d_tmp, sizeof(double) * h->ncell * h->F));
    while (done < nframes) {
        const int nf = (int)std::min<int64_t>(h->F, nframes - done);
        CU(cudaMemcpyAsync(d_tmp, d1 + done * h->ncell, sizeof(double) * h->ncell * nf, cudaMemcpyHostToDevice, h->s_comp));
        const int npairs = (nf + 1) / 2;
        SplatArgs sa{};
        sa.src_density = d_tmp; sa.nframes = nf; sa.vol = h->d_vol; sa.dens_dump = h->d_dump; sa.gp = gp; sa.tt = h->tt;
        sa.zplan = h->ax[2].plan; sa.twz = h->ax[2].d_tw; sa.err_flag = h->d_err;
        sa.tws = h->d_tws; sa.tws_n = h->tws_n; sa.tws_off = h->tws_off;
        cudaError_t ce = mdsf_launch_splat(gp.lcol, SPLAT_DENSITY_, h->native_fft, dim3(gp.ntx * gp.nty, npairs), h->splat_smem, h->s_comp, sa);
        if (ce != cudaSuccess) { cudaFree(d_tmp); return fail(MDSF_ECUDA, "density tile kernel launch failed: %s", cudaGetErrorString(ce)); }
        ++h->launches;
        int rc = transform_and_accumulate(h, nf);
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
    CU(cudaStreamSynchronize(h->s_copy));
    CU(cudaStreamSynchronize(h->s_prep));
    CU(cudaStreamSynchronize(h->s_comp));
    CU(cudaStreamSynchronize(h->s_back));
    CU(cudaMemcpy(h->h_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    const int code = *h->h_err;
    if (code) {
        *h->h_err = 0;
        cudaMemset(h->d_err, 0, sizeof(int));
        if (code == 4) return fail(MDSF_ECUDA, "internal: pair lists exceed their computed capacity");
        if (code == 5) return fail(MDSF_ECUDA, "internal: a dependency of the fused y/x pass never arrived (bounded wait expired)");
        if (code == 2) return fail(MDSF_ERANGE, "fixed-point density accumulator overflow (> 2048 peak amplitudes in one cell)");
        return fail(MDSF_ERANGE, "an atom's Gaussian stamp leaves the padded grid (coordinate more than one box outside the cell, NaN, or half width > Nborder); the reference fails with a numpy shape error here");
    }
    return MDSF_OK;
}

extern "C" int mdsf_export_sf_device(mdsf_handle* h, void* sf_device) {
    if (!h || !sf_device) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const GridParams& gp = h->gp;
    const long long total = (long long)gp.n[0] * gp.n[1] * (gp.n[2] / 2 + 1);
    GridParams ge = gp;
    if (!h->native_fft) { ge.lw = gp.n[2]; ge.nch = 1; }
    export_sf_kernel<<<grid_for(total, 256, h->nsm), 256, 0, h->s_comp>>>(h->d_P, (double*)sf_device, h->ax[0].d_rev, h->ax[1].d_rev, h->ax[2].d_rev, ge);
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

// ------------------------------------------------------------------------------------------ plot grids
// sfplt = get_dplot(sf) (dens.py:142-163) as ONE gather: element (i, j, k) of the cropped view is cell (X, Y, Z) =
// (i+1, j+1, k+1) of the (Nx, Ny, 2M-1) array the reference assembles from four quadrant copies (Z >= M-1:
// sf[(X - Nx/2) mod Nx][(Y - Ny/2) mod Ny][Z - (M-1)]), the point inversion of that half (Z < M-1:
// sf[(Nx/2 - X) mod Nx][(Ny/2 - Y) mod Ny][(M-1) - Z]) and the DC cell replaced by 1/3 (+x + +y + +z neighbours), :161.
__device__ __forceinline__ double dplot_cell(const double* __restrict__ sf, int nx, int ny, int m, int X, int Y, int Z) {
    const int hz = m - 1;
    int sx, sy, sz;
    if (Z >= hz) { sx = X - nx / 2; sy = Y - ny / 2; sz = Z - hz; }
    else { sx = nx / 2 - X; sy = ny / 2 - Y; sz = hz - Z; }
    sx = (sx % nx + nx) % nx; sy = (sy % ny + ny) % ny;
    return sf[((long long)sx * ny + sy) * m + sz];
}
__device__ __forceinline__ double sfplt_value(const double* __restrict__ sf, int nx, int ny, int m, int i, int j, int k) {
    const int X = i + 1, Y = j + 1, Z = k + 1;
    const int cx = nx / 2, cy = ny / 2, cz = (2 * m - 1) / 2;
    if (X == cx && Y == cy && Z == cz)
        return 1.0 / 3.0 * ((dplot_cell(sf, nx, ny, m, cx + 1, cy, cz) + dplot_cell(sf, nx, ny, m, cx, cy + 1, cz)) + dplot_cell(sf, nx, ny, m, cx, cy, cz + 1));
    return dplot_cell(sf, nx, ny, m, X, Y, Z);
}
// out[(i*n1 + j)*n2 + k][0..2] = axis values, [3] = sfplt(i,j,k) (kgridplt, dens.py:336-344) or 0 (kgrid, :325-334);
// channels == 1: plain sfplt
static __global__ void plot_grid_kernel(const double* __restrict__ sf, int nx, int ny, int m, const double* __restrict__ a0,
                                        const double* __restrict__ a1, const double* __restrict__ a2, int n0, int n1, int n2,
                                        int channels, int with_sf, double* __restrict__ out)
{
    const long long total = (long long)n0 * n1 * n2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e % n2);
        const long long t = e / n2;
        const int j = (int)(t % n1), i = (int)(t / n1);
        const double v = with_sf ? sfplt_value(sf, nx, ny, m, i, j, k) : 0.0;
        if (channels == 1) out[e] = v;
        else {
            double2* o = reinterpret_cast<double2*>(out + 4 * e);
            o[0] = make_double2(a0[i], a1[j]);
            o[1] = make_double2(a2[k], v);
        }
    }
}

// sfplt (Nx-2, Ny-2, 2M-3), kgrid (Nx, Ny, M, 4), kgridplt (Nx-2, Ny-2, 2M-3, 4) of dens.py:323-344 assembled on the GPU from
// the resident S(q) and copied to the caller's host arrays (any of the three may be NULL).  The six axis-value vectors are
// the reference's per-index scalar expressions, evaluated by the caller (host, once): kx[Nx], ky[Ny], kz[M] for kgrid and
// px[Nx-2], py[Ny-2], pz[2M-3] for kgridplt -- the GPU only broadcasts, gathers and interleaves, so every value is bit-exact.
extern "C" int mdsf_export_plot_grids(mdsf_handle* h, const double* kx, const double* ky, const double* kz,
                                      const double* px, const double* py, const double* pz,
                                      double* sfplt_host, double* kgrid_host, double* kgridplt_host) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    if ((kgrid_host && (!kx || !ky || !kz)) || (kgridplt_host && (!px || !py || !pz))) return fail(MDSF_EINVAL, "axis values missing");
    int rc = mdsf_sync(h);
    if (rc) return rc;
    rc = mdsf_export_sf_device(h, h->d_sf);
    if (rc) return rc;
    const GridParams& gp = h->gp;
    const int nx = gp.n[0], ny = gp.n[1], m = gp.n[2] / 2 + 1;
    if (nx < 4 || ny < 4 || m < 3) return fail(MDSF_EINVAL, "grid too small for the plot view");
    const int q0 = nx - 2, q1 = ny - 2, q2 = 2 * m - 3;
    const long long nplt = (long long)q0 * q1 * q2, ngrid = (long long)nx * ny * m;
    double *d_ax = nullptr, *d_out = nullptr;
    const size_t nax = (size_t)nx + ny + m + q0 + q1 + q2;
    CU(cudaMalloc(&d_ax, sizeof(double) * nax));
    size_t need = 0;
    if (sfplt_host) need = std::max(need, sizeof(double) * (size_t)nplt);
    if (kgrid_host) need = std::max(need, sizeof(double) * 4 * (size_t)ngrid);
    if (kgridplt_host) need = std::max(need, sizeof(double) * 4 * (size_t)nplt);
    cudaError_t ce = need ? cudaMalloc(&d_out, need) : cudaSuccess;
    if (ce != cudaSuccess) { cudaFree(d_ax); return fail(MDSF_ECUDA, "plot grids: %s", cudaGetErrorString(ce)); }
    double *dkx = d_ax, *dky = dkx + nx, *dkz = dky + ny, *dpx = dkz + m, *dpy = dpx + q0, *dpz = dpy + q1;
    auto up = [&](double* d, const double* s, int n) { return s ? cudaMemcpy(d, s, sizeof(double) * n, cudaMemcpyHostToDevice) : cudaSuccess; };
    ce = up(dkx, kx, nx); if (ce == cudaSuccess) ce = up(dky, ky, ny); if (ce == cudaSuccess) ce = up(dkz, kz, m);
    if (ce == cudaSuccess) ce = up(dpx, px, q0); if (ce == cudaSuccess) ce = up(dpy, py, q1); if (ce == cudaSuccess) ce = up(dpz, pz, q2);
    auto run = [&](int n0, int n1, int n2, const double* a0, const double* a1, const double* a2, int channels, int with_sf, double* host) {
        if (ce != cudaSuccess || !host) return;
        const long long total = (long long)n0 * n1 * n2;
        plot_grid_kernel<<<grid_for(total, 256, h->nsm), 256, 0, h->s_comp>>>(h->d_sf, nx, ny, m, a0, a1, a2, n0, n1, n2, channels, with_sf, d_out);
        ++h->launches;
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->s_comp);
        if (ce == cudaSuccess) ce = cudaMemcpy(host, d_out, sizeof(double) * channels * total, cudaMemcpyDeviceToHost);
    };
    run(q0, q1, q2, dpx, dpy, dpz, 1, 1, sfplt_host);
    run(nx, ny, m, dkx, dky, dkz, 4, 0, kgrid_host);
    run(q0, q1, q2, dpx, dpy, dpz, 4, 1, kgridplt_host);
    cudaFree(d_ax);
    if (d_out) cudaFree(d_out);
    if (ce != cudaSuccess) return fail(MDSF_ECUDA, "plot grids: %s", cudaGetErrorString(ce));
    return MDSF_OK;
}

// Cylindrical average of plot2d.PLOT_RAD_NEW (reference plot2d.py:575-632) over a host structure-factor volume (the
// channel-3 values of kgridplt and its three axis vectors, as plot2d.py:571-574 slices them).  Stand-alone: needs no
// engine handle, like the reference's plot2d which works from the sf npz.  oa_host[rbins][zbins] = mean over theta
// (NaN where a ring leaves the grid; the caller applies the reference's nanmin fill and normalisation, plot2d.py:634-638).
extern "C" int mdsf_cylindrical_average(int device, const double* sf_host, int n0, int n1, int n2, const double* X, const double* Y,
                                        const double* Z, const double* binv9, int rbins, const double* rarr, const int* ntheta,
                                        int zbins, const double* zar, double* oa_host) {
    if (!sf_host || !X || !Y || !Z || !binv9 || !rarr || !ntheta || !zar || !oa_host) return fail(MDSF_EINVAL, "null argument");
    if (n0 < 2 || n1 < 2 || n2 < 2 || rbins < 1 || zbins < 1) return fail(MDSF_EINVAL, "cylindrical average: bad sizes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(MDSF_ECUDA, "no CUDA device available; libmdsf has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(MDSF_EINVAL, "device %d out of range", device);
    CU(cudaSetDevice(device));
    const size_t nsf = (size_t)n0 * n1 * n2;
    double *d_sf = nullptr, *d_ax = nullptr, *d_oa = nullptr;
    int* d_nt = nullptr;
    cudaError_t ce = cudaMalloc(&d_sf, sizeof(double) * nsf);
    const size_t nax = (size_t)n0 + n1 + n2 + 9 + rbins + zbins;
    if (ce == cudaSuccess) ce = cudaMalloc(&d_ax, sizeof(double) * nax);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_oa, sizeof(double) * (size_t)rbins * zbins);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_nt, sizeof(int) * rbins);
    double *dX = d_ax, *dY = dX + n0, *dZ = dY + n1, *dB = dZ + n2, *dR = dB + 9, *dZa = dR + rbins;
    auto up = [&](void* d, const void* s, size_t b) { if (ce == cudaSuccess) ce = cudaMemcpy(d, s, b, cudaMemcpyHostToDevice); };
    up(d_sf, sf_host, sizeof(double) * nsf);
    if (ce == cudaSuccess) { up(dX, X, 8 * (size_t)n0); up(dY, Y, 8 * (size_t)n1); up(dZ, Z, 8 * (size_t)n2); up(dB, binv9, 72); up(dR, rarr, 8 * (size_t)rbins); up(dZa, zar, 8 * (size_t)zbins); up(d_nt, ntheta, 4 * (size_t)rbins); }
    if (ce == cudaSuccess) {
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
        cyl_average_kernel<<<grid_for((long long)rbins * zbins, 128, nsm), 128>>>(d_sf, n0, n1, n2, dX, dY, dZ, dB, rbins, dR, d_nt, zbins, dZa, d_oa);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpy(oa_host, d_oa, sizeof(double) * (size_t)rbins * zbins, cudaMemcpyDeviceToHost);
    cudaFree(d_sf); cudaFree(d_ax); cudaFree(d_oa); cudaFree(d_nt);
    if (ce != cudaSuccess) return fail(MDSF_ECUDA, "cylindrical average: %s", cudaGetErrorString(ce));
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
static int fetch_recs(mdsf_handle* h, int64_t frame, std::vector<AtomRec>& recs) {
    if (frame < 0 || frame >= h->last_batch_frames) return fail(MDSF_EINVAL, "frame %lld not in the last batch (%d frames)", (long long)frame, h->last_batch_frames);
    if (!h->d_type) return fail(MDSF_ESTATE, "no atoms set");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    recs.resize(h->natoms);
    CU(cudaMemcpy(recs.data(), h->d_recs_last + frame * h->natoms, sizeof(AtomRec) * h->natoms, cudaMemcpyDeviceToHost));
    return MDSF_OK;
}

extern "C" int mdsf_debug_cell_indices(mdsf_handle* h, int64_t frame, int32_t* ir_out) {
    if (!h || !ir_out) return fail(MDSF_EINVAL, "null argument");
    std::vector<AtomRec> recs;
    int rc = fetch_recs(h, frame, recs);
    if (rc) return rc;
    for (long long a = 0; a < h->natoms; ++a) for (int d = 0; d < 3; ++d) ir_out[a * 3 + d] = recs[a].ir[d];
    return MDSF_OK;
}

extern "C" int mdsf_debug_coords(mdsf_handle* h, int64_t frame, double* r_out) {
    if (!h || !r_out) return fail(MDSF_EINVAL, "null argument");
    std::vector<AtomRec> recs;
    int rc = fetch_recs(h, frame, recs);
    if (rc) return rc;
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
extern "C" const char* mdsf_splat_path(const mdsf_handle* h) {
    if (!h) return "";
    return h->splat_mode == SPLAT_GENERAL_ ? "register-general" : (h->splat_mode == SPLAT_MONO_ ? "register-mono" : "register-ortho");
}
extern "C" int mdsf_batch_frames(const mdsf_handle* h) { return h ? h->F : 0; }
extern "C" int mdsf_set_pretransform(mdsf_handle* h, int32_t enabled, double sin_theta, double cos_theta) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    if (enabled && !(sin_theta != 0.0)) return fail(MDSF_EINVAL, "sin(theta) must be non-zero");
    h->mono = enabled ? 1 : 0;
    h->mono_sin = sin_theta;
    h->mono_cos = cos_theta;
    return MDSF_OK;
}
extern "C" int mdsf_geometry(const mdsf_handle* h, int32_t* out8) {
    if (!h || !out8) return fail(MDSF_EINVAL, "null argument");
    const GridParams& gp = h->gp;
    out8[0] = 1 << ((gp.lcol + 1) / 2); out8[1] = 1 << (gp.lcol / 2); out8[2] = gp.zw; out8[3] = gp.nslab;
    out8[4] = gp.lw; out8[5] = h->pgy.W; out8[6] = h->pgx.W; out8[7] = (int)(h->splat_smem / 1024);
    return MDSF_OK;
}
extern "C" int mdsf_enable_timing(mdsf_handle* h, int32_t on) {
    if (!h) return fail(MDSF_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    h->timing = on != 0;
    for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
    h->tev.clear();
    return MDSF_OK;
}
extern "C" int mdsf_stage_ms(mdsf_handle* h, double* out6, int64_t* batches) {
    if (!h || !out6) return fail(MDSF_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < 6; ++i) out6[i] = 0;
    const size_t nb = h->tev.size() / 8;
    for (size_t b = 0; b < nb; ++b) {
        // tv: 0 copy start (s_copy), 1 prep start, 2 binned (s_prep), 6 splat start, 3 splat done, 4 y done, 5 end (s_comp)
        cudaEvent_t* tv = &h->tev[b * 8];
        float ms;
        CU(cudaEventElapsedTime(&ms, tv[0], tv[1])); out6[0] += ms;     // copy (overlaps the previous batch's kernels)
        CU(cudaEventElapsedTime(&ms, tv[1], tv[2])); out6[1] += ms;     // prep + bin (overlaps the previous batch)
        CU(cudaEventElapsedTime(&ms, tv[6], tv[3])); out6[2] += ms;
        CU(cudaEventElapsedTime(&ms, tv[3], tv[4])); out6[3] += ms;
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
