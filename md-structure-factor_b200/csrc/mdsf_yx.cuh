// K4+K5 fused: y pass -> x pass + |C|^2 accumulation in ONE persistent kernel; the hand-over between the two passes
// stays in L2.
//
// Layout: z-chunked volumes [pair][z/4][x][y][4] (lw = 4).  For one step t = (z chunk g, pair q) the data of the two
// passes is a 16.8 MB slab at 512^2 (Nx*Ny*4 cells): small enough for three of them plus the chunk's slice of P to
// live in the 126 MB L2.  Work items, handed out in one global order by an atomic counter:
//   Y(t, x):  [Ny][4] tile of volume q at (chunk g, x) -- one contiguous 32 KB run in HBM -- arrives by TMA bulk copies
//             (cp.async.bulk + mbarrier) into padded shared memory, is transformed along y and leaves by bulk stores
//             into ring slot t % 3 of the scratch buffer.  Nobody reads that slot from HBM again.
//   X(t, ky): the [Nx][4] tile at fixed ky of ring slot t % 3 (64-byte rows, cp.async) is transformed along x,
//             |C|^2 is formed in registers and added to P.
// Order: phase p hands out Y(p) then X(p - LAG).  X(t) waits for ydone[t] == Nx, Y(t) for xdone[t - RING] == Ny (ring
// reuse): with LAG = 2, RING = 4 every wait is on items handed out at least 1.5 phases (1.5 (Nx + Ny) tiles) earlier -- more
// than the tiles all resident CTAs hold in flight -- so it practically never spins (LAG = 1 / RING = 3 did: 2.5x slower), and never
// deadlocks whatever the number of resident CTAs (waits only point backwards in the hand-out order).  The adds into P
// of one (chunk, ky) tile are ordered over the pairs by a per-tile counter (pseq), so S(q) stays bitwise
// reproducible; between two pairs the P slice (8.4 MB) is still in L2, so P costs HBM one read-modify-write per
// batch.  HBM traffic per pair: 16 N^3 (read of the z-transformed volume) instead of 48 N^3 for separate y and x passes.
#pragma once
#include "mdsf_common.cuh"
#include "mdsf_fft.cuh"

#ifndef MDSF_YX_THREADS
#define MDSF_YX_THREADS 256        // compute threads per CTA (+ one producer warp): one radix-8 butterfly per thread and stage
#endif
#define MDSF_YX_W 4
#ifndef MDSF_YX_LAG
#define MDSF_YX_LAG 2          // phases the X items run behind the Y items of the same step
#endif
#ifndef MDSF_YX_RING
#define MDSF_YX_RING 4         // scratch slabs in flight: Y(t) reuses the slab X(t - RING) has drained (RING >= LAG + 2)
#endif
#ifndef MDSF_YX_GROUP
#define MDSF_YX_GROUP 2        // tiles per work item (one claim, one dependency wait, one completion signal per item)
#endif

struct YXParams {
    const double2* vol;        // [npairs][nch][Nx][Ny][4]
    double2* scratch;          // [RING][Nx][Ny][4]
    double* P;                 // [nch][Nx][Ny][4]
    const double2* twy;        // per-stage twiddle tables of the y plan (w^(n2 N/L) per stage, as in the z path)
    const double2* twx;
    int nch, npairs;
    unsigned* ctl;             // [0] work counter, then ydone[T], xdone[T], pseq[nch*Ny]
    int* err;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded: a phase that has not completed after ~4 s of wall time is a protocol bug -- trap (the launch fails with an
// error the host reports) instead of spinning forever and wedging the device.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if ((spins & 1023u) == 1023u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ULL) __trap();
        }
    }
}
// L2 policies: the z-transformed volume streams through once (evict first), the scratch ring is written and read
// back within a few phases and must not be pushed out by that stream (evict last)
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes, unsigned long long policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one thread waits until *flag >= want (bounded: a dependency that never arrives sets the error flag instead of hanging)
__device__ __forceinline__ bool spin_until(const unsigned* flag, unsigned want, int* err) {
    for (long long spins = 0; spins < (1LL << 26); ++spins) {
        if (ld_acquire(flag) >= want) return true;
        __nanosleep(64);
    }
    atomicExch(err, 5);
    return false;
}

// row r of a [N][4] tile sits at padded row r + (r >> 3): the stride-8 stage then alternates the two 64-byte halves
// of a 128-byte bank line (conflict-free quarter-warps); rows of 64 bytes
__device__ __forceinline__ int yx_row(int r) { return r + (r >> 3); }

// in-place DIF stage over the rows of a padded [N][4] double2 tile; LAST: outputs stay in registers and go to `sink`
template <int N, int R, int L, int TOFF, bool LAST, typename Sink>
__device__ __forceinline__ void yx_stage(double2* __restrict__ tile, const double2* __restrict__ tws, Sink&& sink)
{
    constexpr int M = L / R, ITEMS = (N / R) * MDSF_YX_W;
    for (int it = threadIdx.x; it < ITEMS; it += MDSF_YX_THREADS) {
        const int f = it & (MDSF_YX_W - 1), bf = it >> 2;
        const int b = bf / M, n2 = bf % M;
        const int base = b * L + n2;
        double xr[R], xi[R];
        int a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            a[j] = yx_row(base + j * M) * MDSF_YX_W + f;
            const double2 v = tile[a[j]];
            xr[j] = v.x; xi[j] = v.y;
        }
        Dft<R>::run(xr, xi, nullptr, nullptr, N);
        if (M > 1) {
            const double2 w1 = __ldg(tws + TOFF + n2);
            double wr[R], wi[R];
            wr[1] = w1.x; wi[1] = w1.y;
#pragma unroll
            for (int k = 2; k < R; ++k) {
                const int ka = k >> 1, kb = k - ka;
                wr[k] = wr[ka] * wr[kb] - wi[ka] * wi[kb];
                wi[k] = wr[ka] * wi[kb] + wi[ka] * wr[kb];
            }
#pragma unroll
            for (int k = 1; k < R; ++k) {
                const double yr = xr[k] * wr[k] - xi[k] * wi[k];
                xi[k] = xr[k] * wi[k] + xi[k] * wr[k];
                xr[k] = yr;
            }
        }
        if (LAST) {
#pragma unroll
            for (int k = 0; k < R; ++k) sink(base + k * M, f, xr[k], xi[k]);
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) tile[a[k]] = make_double2(xr[k], xi[k]);
        }
    }
}

// all stages of an N-point axis (N = 256: 4*8*8, N = 512: 8*8*8 -- the radix lists factorize() builds with radices <= 8)
template <int N, typename Sink>
__device__ __forceinline__ void yx_fft_unused(double2* tile, const double2* tws, Sink&& sink) {
    auto none = [](int, int, double, double) {};
    if (N == 512) {
        yx_stage<N, 8, 512, 0, false>(tile, tws, none);
        __syncthreads();
        yx_stage<N, 8, 64, 64, false>(tile, tws, none);
        __syncthreads();
        yx_stage<N, 8, 8, 72, true>(tile, tws, sink);
    } else {
        yx_stage<N, 4, 256, 0, false>(tile, tws, none);
        __syncthreads();
        yx_stage<N, 8, 64, 64, false>(tile, tws, none);
        __syncthreads();
        yx_stage<N, 8, 8, 72, true>(tile, tws, sink);
    }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// arrive on `bar` once all cp.async copies this thread issued so far have landed (counted in the barrier's expected arrivals)
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the compute threads (warps 0 .. MDSF_YX_THREADS/32 - 1); the producer warp never joins it
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(MDSF_YX_THREADS) : "memory"); }
__device__ __forceinline__ bool bar_compute_or(bool pred) {
    unsigned r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "barrier.red.or.pred q, 1, %2, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n" : "=r"(r) : "r"((unsigned)pred), "n"(MDSF_YX_THREADS) : "memory");
    return r != 0;
}

#define MDSF_YX_CTA (MDSF_YX_THREADS + 32)        // compute warps + 1 producer warp

// Warp-specialised, double-buffered: warp 4 claims the next item, waits for its dependency and starts its loads (TMA
// bulk copies for Y tiles, cp.async for the strided X tiles) into the other tile buffer while warps 0-3 transform the
// current one.  Completion signals (ydone / xdone / pseq) of an item are sent one item later, right before the last
// stage of the next item -- the bulk stores and the fence behind them have then long finished -- or at once when the
// compute warps would otherwise block (a withheld signal must never be what somebody, possibly our own producer,
// waits for).
template <int NY, int NX>
__global__ void __launch_bounds__(MDSF_YX_CTA)
yx_pass_kernel(YXParams p)
{
    constexpr int W = MDSF_YX_W;
    constexpr int NMAX = NY > NX ? NY : NX;
    constexpr int TILE = (NMAX + NMAX / 8) * W;                // cells of one padded tile buffer
    extern __shared__ double2 yx_smem[];
    __shared__ unsigned long long full[2], empty[2];
    __shared__ int4 desc[2];                                   // x: 1 = Y item, 0 = X item, -1 = end; y: step t; z: x / ky
    const int T = p.nch * p.npairs;                            // steps
    unsigned* work = p.ctl;
    unsigned* ydone = p.ctl + 1;
    unsigned* xdone = ydone + T;
    unsigned* pseq = xdone + T;
    const unsigned per_phase = NX + NY, total = (unsigned)(T + 1) * per_phase;
    const long long chunk_cells = (long long)NX * NY * W;      // cells of one (pair, chunk) slab = of one ring slot
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 32); mbar_init(&full[1], 32);
        mbar_init(&empty[0], MDSF_YX_THREADS); mbar_init(&empty[1], MDSF_YX_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    constexpr int GRP = MDSF_YX_GROUP;
    const unsigned per_phase_i = (NX + NY) / GRP, total_i = (unsigned)(T + MDSF_YX_LAG) * per_phase_i;      // in items of GRP tiles
    const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    if (warp == MDSF_YX_THREADS / 32) {
        // ------------------------------------------------------------------ producer
        int n = 0;                                             // tile buffers filled so far
        for (;;) {
            int is_y = 0, t = -1, r = 0;
            bool end = false;
            for (;;) {                                         // claim; skip the empty halves of the first / last phase
                unsigned slot = 0;
                if (lane == 0) slot = atomicAdd(work, 1u);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot >= total_i) { end = true; break; }
                const int phase = (int)(slot / per_phase_i);
                r = (int)(slot - (unsigned)phase * per_phase_i) * GRP;       // first x (Y item) / NX + first ky (X item)
                is_y = r < NX;
                t = is_y ? phase : phase - MDSF_YX_LAG;
                if (t >= 0 && t < T) break;
            }
            int ok = 1;
            if (!end && lane == 0) {
                if (is_y) ok = (t < MDSF_YX_RING) ? 1 : (spin_until(xdone + (t - MDSF_YX_RING), NY, p.err) ? 1 : 0);   // ring slot t % 3 free
                else ok = spin_until(ydone + t, NX, p.err) ? 1 : 0;
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (end || !ok) {
                const int buf = n & 1, use = n >> 1;
                mbar_wait(&empty[buf], (use & 1) ^ 1);
                if (lane == 0) desc[buf] = make_int4(-1, 0, 0, 0);
                mbar_arrive(&full[buf]);
                break;
            }
            const int g = t / p.npairs, q = t - g * p.npairs;
            double2* ring = p.scratch + (long long)(t % MDSF_YX_RING) * chunk_cells;
            for (int sub = 0; sub < GRP; ++sub, ++n) {
                const int buf = n & 1, use = n >> 1;
                double2* tile = yx_smem + (size_t)buf * TILE;
                mbar_wait(&empty[buf], (use & 1) ^ 1);         // the compute warps are done with this buffer
                if (is_y) {
                    const int x = r + sub;
                    if (lane == 0) {
                        desc[buf] = make_int4(1, t, x, sub == GRP - 1);
                        fence_async_smem();
                        mbar_expect_tx(&full[buf], NY * W * 16);   // counts as lane 0's arrival
                    } else {
                        mbar_arrive(&full[buf]);
                    }
                    __syncwarp();                              // expect_tx precedes every complete_tx
                    const double2* src = p.vol + ((long long)q * p.nch + g) * chunk_cells + (long long)x * NY * W;
                    for (int k = lane; k < NY / 8; k += 32)    // 8 rows (512 B) per bulk copy, into the padded row positions
                        bulk_g2s(tile + (size_t)(9 * k) * W, src + (size_t)(8 * k) * W, 512, &full[buf], pol_stream);
                } else {
                    const int ky = r - NX + sub;
                    if (lane == 0) desc[buf] = make_int4(0, t, ky, sub == GRP - 1);
                    // [NX][4] tile: 64-byte rows at stride NY*4 cells, through L2 only (cp.async.cg)
                    const double2* src = ring + (long long)ky * W;
                    for (int i = lane; i < NX * W; i += 32) {
                        const int row = i >> 2, f = i & 3;
                        const unsigned d = smem_u32(tile + yx_row(row) * W + f);
                        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(src + (long long)row * NY * W + f), "l"(pol_keep) : "memory");
                    }
                    cp_async_arrive_noinc(&full[buf]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- compute warps
    int pend_kind = -1, pend_t = 0, pend_idx = 0;              // item whose completion signal is still owed (-1: none)
    auto flush = [&]() {
        if (pend_kind < 0) return;
        if (pend_kind == 1) bulk_wait0();                      // my bulk stores of that item's Y tiles are complete
        bar_compute();
        if (threadIdx.x == 0) {
            __threadfence();                                   // cumulative: orders the stores of all 128 threads (seen through the barrier)
            if (pend_kind == 1) {
                atomicAdd(ydone + pend_t, (unsigned)GRP);
            } else {
                const int g = pend_t / p.npairs, q = pend_t - g * p.npairs;
                atomicExch(pseq + (g * (NY / GRP) + pend_idx / GRP), (unsigned)(q + 1));
                atomicAdd(xdone + pend_t, (unsigned)GRP);
            }
        }
        pend_kind = -1;
    };
    auto none = [](int, int, double, double) {};
    bool pseq_ok = false;                                      // this X item's turn on P has been awaited
    for (int n = 0;; ++n) {
        const int buf = n & 1, use = n >> 1;
        double2* tile = yx_smem + (size_t)buf * TILE;
        if (bar_compute_or(!mbar_test(&full[buf], use & 1))) { flush(); mbar_wait(&full[buf], use & 1); }
        const int4 d = desc[buf];
        if (d.x < 0) break;
        const int t = d.y, g = t / p.npairs, q = t - g * p.npairs;
        if (d.x == 1) {
            constexpr int N = NY;
            if (N == 512) yx_stage<N, 8, 512, 0, false>(tile, p.twy, none); else yx_stage<N, 4, 256, 0, false>(tile, p.twy, none);
            bar_compute();
            yx_stage<N, 8, 64, 64, false>(tile, p.twy, none);
            bar_compute();
            flush();
            auto to_tile = [&](int row, int f, double re, double im) { tile[yx_row(row) * W + f] = make_double2(re, im); };
            yx_stage<N, 8, 8, 72, true>(tile, p.twy, to_tile);
            fence_async_smem();                                // my generic-proxy writes, before anybody's bulk store reads them
            bar_compute();
            double2* dst = p.scratch + (long long)(t % MDSF_YX_RING) * chunk_cells + (long long)d.z * NY * W;
            for (int k = threadIdx.x; k < NY / 8; k += MDSF_YX_THREADS)
                bulk_s2g(dst + (size_t)(8 * k) * W, tile + (size_t)(9 * k) * W, 512, pol_keep);
            bulk_commit();
            bulk_wait_read0();                                 // the stores have read the tile: the producer may refill it
            mbar_arrive(&empty[buf]);
            if (d.w) { pend_kind = 1; pend_t = t; pend_idx = d.z; }
        } else {
            constexpr int N = NX;
            const int ky = d.z;
            if (N == 512) yx_stage<N, 8, 512, 0, false>(tile, p.twx, none); else yx_stage<N, 4, 256, 0, false>(tile, p.twx, none);
            bar_compute();
            yx_stage<N, 8, 64, 64, false>(tile, p.twx, none);
            bar_compute();
            flush();
            // the adds of pair q into this group of (chunk, ky) tiles of P follow those of pair q - 1
            double* Pt = p.P + (long long)g * chunk_cells + (long long)ky * W;
            bool dead = false;
            auto to_P = [&](int row, int f, double re, double im) {
                if (!pseq_ok) {                                // first output of this thread for this item: pair q - 1 must be in
                    if (q > 0 && !spin_until(pseq + (g * (NY / GRP) + ky / GRP), (unsigned)q, p.err)) dead = true;
                    pseq_ok = true;
                }
                if (dead) return;
                double* cell = Pt + (long long)row * NY * W + f;
                __stcg(cell, __ldcg(cell) + (re * re + im * im));
            };
            yx_stage<N, 8, 8, 72, true>(tile, p.twx, to_P);
            mbar_arrive(&empty[buf]);                          // (stage 3 only read my own cells of the tile)
            if (d.w) { pend_kind = 0; pend_t = t; pend_idx = ky; pseq_ok = false; }
        }
    }
    flush();
}
