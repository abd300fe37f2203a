// K4 / K5 for 512-point axes: persistent, warp-specialised y and x passes fed by the TMA copy engine.
//
// Layout: z-chunked volumes [pair][z/8][x][y][8] (lw = 8).  A tile is [512][8] complex = 64 KB:
//   y pass: tile (pair, chunk, x) is ONE contiguous 64 KB run -> 32 bulk copies of 2 KB in, 32 bulk stores of 2 KB out;
//   x pass: tile (pair, chunk, y) is 512 rows of 128 bytes, 64 KB apart.  128-byte bulk copies are far too slow for that
//           (measured: 27 ms per 16 c3 frames, the copy engine retires one small copy per ~60 cycles); every compute thread
//           issues eight 16-byte cp.async copies instead (8 threads = one row), two tiles ahead, completion counted on the
//           buffer's mbarrier (cp.async.mbarrier.arrive.noinc).  Tried and dropped: a warp that requests the contiguous runs the
//           G CTAs' pieces form per x row into L2 ahead of the copies (cp.async.bulk.prefetch.L2) -- DRAM reads doubled
//           (35.4 GB instead of 18.3 GB per 16 frames, 11.2 ms): the CTAs drift apart and the lines are gone before use;
//           clusters of 2 / 4 CTAs kept in step by a cluster barrier per tile (8.2 / 16.8 ms); the .L2::256B hint (no change).
//           The pass is bound by its arithmetic (three barrier-separated radix-8 stages, ~3 us per tile and SM), not by DRAM:
//           4.22 ms per 16 c3 frames = 70 % of the HBM peak once the tile loop carried pointers instead of 64-bit divisions.
// One CTA per SM walks its tiles (unit u = blockIdx.x + k * gridDim.x) through a ring of three 64 KB buffers:
//   producer warp:  [y: wait until the compute warps released the buffer, store it (cp.async.bulk shared -> global), wait
//                   until the store has read it]  ->  arm the buffer's mbarrier with the tile's byte count  ->  issue the
//                   bulk loads (cp.async.bulk global -> shared, completion on the mbarrier)
//   16 compute warps: wait on the mbarrier, run the three radix-8 stages in place (one butterfly per thread and stage,
//                   rows are 128 bytes = all 32 banks: every 16-byte access of a quarter-warp is conflict-free without
//                   padding, so the TMA image is the compute layout), release the buffer.
// While one buffer is transformed, one is being loaded and one stored: the loads of a tile never wait for the arithmetic
// of another, which is what kept the register-staged passes (fft3_pass_kernel: load phase / compute phase / store phase
// of two CTAs per SM) at 55-68 % of the HBM peak.  The x pass keeps sum_q |C_q|^2 of its 8 outputs per thread in
// registers over the pairs of the batch and makes one read-modify-write of P per tile (P values requested one pair
// early).  Stage order, twiddles and output positions are those of fft3_pass_kernel<8,8,8> (in place, digit-reversed).
#pragma once
#include "mdsf_yx.cuh"

#define MDSF_TP_THREADS 512                        // compute threads: one radix-8 butterfly each per stage of a [512][8] tile
#define MDSF_TP_CTA (MDSF_TP_THREADS + 32)         // + the producer warp
#define MDSF_TP_NBUF 3
#define MDSF_TP_N 512
#define MDSF_TP_TILE (MDSF_TP_N * 8)               // cells of one tile
#define MDSF_TP_NTW 72                             // twiddle rows: 64 (stage 1) + 8 (stage 2), 7 powers each
#define MDSF_TP_SMEM (MDSF_TP_NBUF * MDSF_TP_TILE * 16 + MDSF_TP_NTW * 7 * 16)

struct TPParams {
    double2* vol;              // [npairs][nch][Nx][Ny][8]
    double* P;                 // [nch][Nx][Ny][8]
    const double2* tws;        // per-stage twiddle tables of the axis (stage_tables: w_N^(n2 N/L))
    int nx, ny, nch, npairs;
};

__device__ __forceinline__ void tp_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MDSF_TP_THREADS) : "memory"); }

// this thread's butterfly of the stage with block length L: rows b*L + n2 + j*M (M = L/8) of column f
template <int L>
__device__ __forceinline__ void tp_load_butterfly(const double2* __restrict__ tile, int (&a)[8], double (&xr)[8], double (&xi)[8]) {
    constexpr int M = L / 8;
    const int f = threadIdx.x & 7, bf = threadIdx.x >> 3;
    const int base = (bf / M) * L + (bf % M);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = (base + j * M) * 8 + f;
        const double2 v = tile[a[j]];
        xr[j] = v.x; xi[j] = v.y;
    }
}
// w^1 .. w^7 of one table entry, by the multiplication tree w^k = w^(k/2) * w^(k - k/2) (what every butterfly used to do
// for itself: 24 of its ~108 fp64 instructions; the passes are bound by exactly those)
__device__ __forceinline__ void tp_powers(double2 w1, double2* __restrict__ out7) {
    double wr[8], wi[8];
    wr[1] = w1.x; wi[1] = w1.y;
#pragma unroll
    for (int k = 2; k < 8; ++k) {
        const int ka = k >> 1, kb = k - ka;
        wr[k] = wr[ka] * wr[kb] - wi[ka] * wi[kb];
        wi[k] = wr[ka] * wi[kb] + wi[ka] * wr[kb];
    }
#pragma unroll
    for (int k = 1; k < 8; ++k) out7[k - 1] = make_double2(wr[k], wi[k]);
}
// x pass: powers in registers per butterfly (its stage loads already wait on shared memory: the table made it slower,
// 4.24 -> 5.26 ms); y pass: table (5.85 -> 5.63 ms = 93 % of the HBM peak)
__device__ __forceinline__ void tp_twiddle_w1(double2 w1, double (&xr)[8], double (&xi)[8]) {
    double wr[8], wi[8];
    wr[1] = w1.x; wi[1] = w1.y;
#pragma unroll
    for (int k = 2; k < 8; ++k) {
        const int ka = k >> 1, kb = k - ka;
        wr[k] = wr[ka] * wr[kb] - wi[ka] * wi[kb];
        wi[k] = wr[ka] * wi[kb] + wi[ka] * wr[kb];
    }
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const double yr = xr[k] * wr[k] - xi[k] * wi[k];
        xi[k] = xr[k] * wi[k] + xi[k] * wr[k];
        xr[k] = yr;
    }
}
__device__ __forceinline__ void tp_twiddle(const double2* __restrict__ w7, double (&xr)[8], double (&xi)[8]) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const double2 w = w7[k - 1];
        const double yr = xr[k] * w.x - xi[k] * w.y;
        xi[k] = xr[k] * w.y + xi[k] * w.x;
        xr[k] = yr;
    }
}

template <bool XPASS>
__global__ void __launch_bounds__(MDSF_TP_CTA, 1)
tma_pass_kernel(TPParams p)
{
    constexpr int N = MDSF_TP_N, TILE = MDSF_TP_TILE, NBUF = MDSF_TP_NBUF;
    extern __shared__ __align__(128) double2 tp_smem[];          // [NBUF][N][8]
    __shared__ unsigned long long full[NBUF], rel[NBUF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // y pass: unit = tile (pair, chunk, x) = 64 KB run number u; x pass: unit = (chunk, y), its npairs tiles in turn
    const long long nunits = XPASS ? (long long)p.nch * p.ny : (long long)p.npairs * p.nch * p.nx;
    const long long mine = nunits > blockIdx.x ? (nunits - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long ntile = XPASS ? mine * p.npairs : mine;      // tiles this CTA moves
    const long long row_stride = (long long)p.ny * 8;            // cells between x rows
    double2* twp = tp_smem + (size_t)NBUF * TILE;                 // [MDSF_TP_NTW][7] twiddle powers, built once per CTA
    if (!XPASS && threadIdx.x < MDSF_TP_NTW) tp_powers(__ldg(p.tws + threadIdx.x), twp + threadIdx.x * 7);
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) { mbar_init(&full[b], XPASS ? MDSF_TP_THREADS : 1); mbar_init(&rel[b], MDSF_TP_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == MDSF_TP_THREADS / 32) {
        // ------------------------------------------------------------------ producer warp (y pass only)
        if (XPASS) return;
        const unsigned long long pol = policy_evict_first();
        for (long long i = 0; i < ntile + NBUF; ++i) {
            const int b = (int)(i % NBUF);
            const unsigned use = (unsigned)(i / NBUF);
            double2* tile = tp_smem + (size_t)b * TILE;
            if (i >= NBUF) {
                mbar_wait(&rel[b], (use - 1) & 1);                   // tile i - NBUF is transformed
                const long long uo = blockIdx.x + (i - NBUF) * (long long)gridDim.x;
                bulk_s2g(p.vol + uo * TILE + lane * (TILE / 32), tile + lane * (TILE / 32), (TILE / 32) * 16, pol);
                bulk_commit();
                bulk_wait_read0();                                     // the store has read the buffer
                __syncwarp();
            }
            if (i >= ntile) continue;
            if (lane == 0) mbar_expect_tx(&full[b], TILE * 16);
            __syncwarp();                                              // expect_tx precedes every complete_tx
            const long long u = blockIdx.x + i * (long long)gridDim.x;
            bulk_g2s(tile + lane * (TILE / 32), p.vol + u * TILE + lane * (TILE / 32), (TILE / 32) * 16, &full[b], pol);
        }
        bulk_wait0();
        return;
    }

    // ---------------------------------------------------------------------- compute warps
    const int f = threadIdx.x & 7, bf = threadIdx.x >> 3;
    const double2* w1 = twp + bf * 7;                        // stage 1: L = 512, M = 64, n2 = bf
    const double2* w2 = twp + (64 + (bf & 7)) * 7;           // stage 2: L = 64, M = 8, n2 = bf % 8
    const double2 w1r = __ldg(p.tws + bf), w2r = __ldg(p.tws + 64 + (bf & 7));      // (x pass: the same entries in registers)
    double acc[8];
    double pv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] = 0.0; pv[k] = 0.0; }
    const unsigned long long polx = policy_evict_first();
    // x pass bookkeeping without divisions in the tile loop (64-bit divides were half of this kernel's instructions): the
    // load stream (two tiles ahead) and the compute stream each carry (pair q, ring buffer b, source / P pointer) and step them.
    const long long pair_cells = (long long)p.nch * p.nx * row_stride;       // cells of one pair volume
    auto group_offset = [&](long long k) {                       // cell offset of (chunk, y) group blockIdx.x + k * gridDim.x, x = 0
        const long long g = blockIdx.x + k * (long long)gridDim.x;
        const int ch = (int)(g / p.ny), y = (int)(g - (long long)ch * p.ny);
        return ((long long)ch * p.nx) * row_stride + (long long)y * 8;
    };
    long long li = 0, lk = 0;                                    // load stream: tile index, group round
    int lq = 0, lb = 0;
    const double2* lsrc = p.vol + group_offset(0) + (long long)bf * row_stride + f;
    auto issue_x = [&]() {                                       // my 8 x 16 bytes of the next x-pass tile -> its ring buffer
        if (li >= ntile) return;
        const unsigned dst = smem_u32(tp_smem + (size_t)lb * TILE + bf * 8 + f);
        const double2* src = lsrc + (long long)lq * pair_cells;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + (unsigned)(j * 64 * 8 * 16)),
                         "l"(src + (long long)(j * 64) * row_stride), "l"(polx) : "memory");
        cp_async_arrive_noinc(&full[lb]);
        ++li;
        lb = (lb + 1 == NBUF) ? 0 : lb + 1;
        if (++lq == p.npairs) { lq = 0; ++lk; lsrc = p.vol + group_offset(lk) + (long long)bf * row_stride + f; }
    };
    if (XPASS) { issue_x(); issue_x(); }
    int b = 0, cq = 0;                                           // compute stream: ring buffer, pair
    unsigned use = 0;
    long long ck = 0;
    double* Pt = XPASS ? p.P + group_offset(0) + f : nullptr;
    for (long long i = 0; i < ntile; ++i) {
        double2* tile = tp_smem + (size_t)b * TILE;
        const int q = cq;
        if (XPASS && q == p.npairs - 1) {                        // my 8 cells of P, needed after this tile's last stage
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) pv[k2] = __ldcs(Pt + (long long)(bf * 8 + k2) * row_stride);
        }
        mbar_wait(&full[b], use & 1);
        int a[8];
        double xr[8], xi[8];
        tp_load_butterfly<512>(tile, a, xr, xi);
        dft8(xr, xi);
        if (XPASS) tp_twiddle_w1(w1r, xr, xi); else tp_twiddle(w1, xr, xi);
#pragma unroll
        for (int k = 0; k < 8; ++k) tile[a[k]] = make_double2(xr[k], xi[k]);
        tp_bar();
        if (XPASS) issue_x();                                    // tile i + 2 (everybody is past the last stage of tile i - 1: its buffer is free)
        tp_load_butterfly<64>(tile, a, xr, xi);
        dft8(xr, xi);
        if (XPASS) tp_twiddle_w1(w2r, xr, xi); else tp_twiddle(w2, xr, xi);
#pragma unroll
        for (int k = 0; k < 8; ++k) tile[a[k]] = make_double2(xr[k], xi[k]);
        tp_bar();
        tp_load_butterfly<8>(tile, a, xr, xi);
        dft8(xr, xi);
        if (!XPASS) {
#pragma unroll
            for (int k = 0; k < 8; ++k) tile[a[k]] = make_double2(xr[k], xi[k]);
            fence_async_smem();                                  // my writes, before the producer's bulk store reads them
            mbar_arrive(&rel[b]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += xr[k] * xr[k] + xi[k] * xi[k];
            if (q == p.npairs - 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    __stcs(Pt + (long long)(bf * 8 + k) * row_stride, pv[k] + acc[k]);
                    acc[k] = 0.0;
                }
            }
        }
        b = (b + 1 == NBUF) ? 0 : b + 1;
        if (b == 0) ++use;
        if (XPASS && ++cq == p.npairs) { cq = 0; ++ck; Pt = p.P + group_offset(ck) + f; }
    }
}
