// K3 (sparse-stamp regime): slab-pipelined fixed-point scatter splat + z pass.
//
// Replaces the per-atom Python loop dens.py:283-308 and the fold dens.py:86-108 when stamps are
// small (a few cells wide, dr >~ 1 A: every BASELINE config except the dr~0.3 A one).  There the
// density is sparse (c2: 0.4 Gaussian terms per cell) and the work is bookkeeping, not arithmetic,
// so the design minimises bookkeeping:
//   * the volume is processed in SLABS of X consecutive x planes (all y, all z, all frame pairs of
//     the batch), small enough (tens of MB) to stay resident in the 126 MB L2;
//   * atoms are binned by slab with a counting sort (K2s) -- order inside a bin is irrelevant, see below;
//   * slab_pipeline_kernel / scatter_chunk: one thread per (atom image, column of its stamp): EX[i]*EY[j]*C[i][j]*EZ[k]
//     from the per-atom factor tables of K1, converted to 64-bit FIXED POINT and added with integer
//     `red.global.add.u64` into the slab accumulator.  Integer addition is associative and
//     commutative, so the result is bitwise deterministic whatever the order of the adds -- the grid
//     is not built from float atomics.  The periodic fold (incl. the corner rule of dens.py:107) is an
//     index map;
//   * slab_pipeline_kernel / zpass_item: reads the slab accumulator (L2 hits), clears it, converts to fp64, runs the z
//     FFT in shared memory and writes the complex pair volume: the density never travels to HBM.
// Fixed point: LSB = 2^(e-52) with 2^e >= max_t Nel/sigma^3, i.e. 2^-52 of the largest peak amplitude
// (finer than an fp64 ulp of any cell near a peak); a cell can hold 2048 peak amplitudes; overflow is
// detected in the z pass and reported.  Measured L2 throughput of the integer reductions on B200:
// 210 G adds/s into a <= 64 MB buffer (tools/micro/atom_bench.cu), 73 G adds/s once the buffer spills.
#pragma once
#include "mdsf_common.cuh"
#include <cooperative_groups.h>
#include "mdsf_fft.cuh"

struct SlabParams {
    int X;               // x planes per slab
    int nslabs;
    double scale;        // 2^(52-e): value -> fixed point
    double inv_scale;    // 2^(e-52)
};

// entry of a slab bin: (frame*natoms + atom) | (sx+1) << 30
#define MDSF_ENTRY_BITS 30

// ---- K2s: counting sort of atom images by slab ------------------------------------------------
// pass 0 counts, pass 1 fills (cursor = exclusive scan of the counts, advanced atomically)
template <int PASS>
__global__ void __launch_bounds__(256)
bin_slabs_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ valid, unsigned* __restrict__ count_or_cursor,
                 unsigned* __restrict__ entries, GridParams gp, TypeTable tt, SlabParams sp, int nframes)
{
    // One block handles a contiguous range of (frame, atom) records.  Slab counters are first
    // accumulated in shared memory; one global atomic per (block, slab) then reserves the block's
    // range, so the hot global counters see gridDim.x adds instead of one per atom image.
    extern __shared__ unsigned s_cnt[];            // [nslabs] counts, then [nslabs] block bases (PASS 1)
    unsigned* s_base = s_cnt + sp.nslabs;
    const long long total = (long long)nframes * gp.natoms;
    const long long per_block = (total + gridDim.x - 1) / gridDim.x;
    const long long lo = blockIdx.x * per_block, hi = min(total, lo + per_block);
    for (int i = threadIdx.x; i < sp.nslabs; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (long long idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
        if (valid[idx] == 0) continue;
        const AtomRec rec = recs[idx];
        const int Ax = tt.halfw[rec.type * 3];
        for (int sx = -1; sx <= 1; ++sx) {
            int xlo, xhi;
            stamp_segment(rec.ir[0], Ax, gp.n[0], sx, xlo, xhi);
            if (xhi <= xlo) continue;
            const int s0 = (xlo - sx * gp.n[0]) / sp.X, s1 = (xhi - 1 - sx * gp.n[0]) / sp.X;
            for (int s = s0; s <= s1; ++s) atomicAdd(&s_cnt[s], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < sp.nslabs; i += blockDim.x) {
        const unsigned c = s_cnt[i];
        if (c) { const unsigned base = atomicAdd(&count_or_cursor[i], c); if (PASS == 1) s_base[i] = base; }
        if (PASS == 1) s_cnt[i] = 0;
    }
    if (PASS == 0) return;
    __syncthreads();
    for (long long idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
        if (valid[idx] == 0) continue;
        const AtomRec rec = recs[idx];
        const int Ax = tt.halfw[rec.type * 3];
        for (int sx = -1; sx <= 1; ++sx) {
            int xlo, xhi;
            stamp_segment(rec.ir[0], Ax, gp.n[0], sx, xlo, xhi);
            if (xhi <= xlo) continue;
            const int s0 = (xlo - sx * gp.n[0]) / sp.X, s1 = (xhi - 1 - sx * gp.n[0]) / sp.X;
            for (int s = s0; s <= s1; ++s)
                entries[s_base[s] + atomicAdd(&s_cnt[s], 1u)] = (unsigned)idx | ((unsigned)(sx + 1) << MDSF_ENTRY_BITS);
        }
    }
}

// exclusive scan of nslabs counts (nslabs <= 4096) -> start[0..nslabs], cursor[0..nslabs) = start;
// also lays out the work queue of the slab pipeline: step t holds the scatter chunks of slab t followed
// by the z-pass items of slab t-1; step_start[t] = first item of step t (t = 0..nslabs+1)
__global__ void __launch_bounds__(1024)
scan_slabs_kernel(const unsigned* __restrict__ count, unsigned* __restrict__ start, unsigned* __restrict__ cursor, int nslabs,
                  unsigned* __restrict__ step_start, int chunk_entries, int X, int nx, int ny, int ncol, int npairs)
{
    __shared__ unsigned part[1024];
    const int per = (nslabs + 1023) / 1024;
    unsigned local = 0;
    for (int i = 0; i < per; ++i) { const int s = threadIdx.x * per + i; if (s < nslabs) local += count[s]; }
    part[threadIdx.x] = local;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned v = threadIdx.x >= d ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - local;
    for (int i = 0; i < per; ++i) {
        const int s = threadIdx.x * per + i;
        if (s < nslabs) { start[s] = run; cursor[s] = run; run += count[s]; }
    }
    if (threadIdx.x == 1023) start[nslabs] = part[1023];
    __syncthreads();
    // items per step (t = 0..nslabs): A(t) + B(t-1); same block-wide scan again
    unsigned items = 0;
    for (int i = 0; i < per + 1; ++i) {
        const int t = threadIdx.x * (per + 1) + i;
        if (t <= nslabs) {
            if (t < nslabs) items += (count[t] + chunk_entries - 1) / chunk_entries;
            if (t >= 1) { const int xc = min(X, nx - (t - 1) * X); items += (unsigned)(((long long)xc * ny + ncol - 1) / ncol) * npairs; }
        }
    }
    part[threadIdx.x] = items;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned v = threadIdx.x >= d ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    run = part[threadIdx.x] - items;
    for (int i = 0; i < per + 1; ++i) {
        const int t = threadIdx.x * (per + 1) + i;
        if (t <= nslabs) {
            step_start[t] = run;
            if (t < nslabs) run += (count[t] + chunk_entries - 1) / chunk_entries;
            if (t >= 1) { const int xc = min(X, nx - (t - 1) * X); run += (unsigned)(((long long)xc * ny + ncol - 1) / ncol) * npairs; }
        }
    }
    if (threadIdx.x == 1023) step_start[nslabs + 1] = part[1023];
}

// ---- K3s: scatter one chunk of <= 64 atom images of a slab into its fixed-point accumulator ------
// acc layout: [pair][x local][y][z][part] int64 (a complex128-shaped cell: re = frame 2q, im = frame 2q+1)
#define MDSF_SC_ENTRIES 256         // atom images per scatter work item (one per thread in the load phase)
struct ScatterSmem {
    int off[MDSF_SC_ENTRIES + 1];
    int4 a[MDSF_SC_ENTRIES];     // x: tbase  y: type | sx+1 << 16 | part << 18   z: pair  w: nrows | i_first << 10 | xl_first << 20
    int4 b[MDSF_SC_ENTRIES];     // x: ir_y - Ay   y: ir_z - Az   z: 2Ay   w: 2Az
    double r[MDSF_SC_ENTRIES * 3];
    int wsum[32];
};

__device__ __forceinline__ void scatter_chunk(ScatterSmem& sm, const AtomRec* __restrict__ recs, const unsigned* __restrict__ entries,
                                              unsigned cb, int ne, const double* __restrict__ atom_tables,
                                              unsigned long long* __restrict__ acc, const GridParams& gp, const TypeTable& tt,
                                              const SlabParams& sp, int slab)
{
    const int X0 = slab * sp.X;
    const int nx = gp.n[0], ny = gp.n[1], nz = gp.n[2];
    if (threadIdx.x < ne) {
        const unsigned ent = entries[cb + threadIdx.x];
        const unsigned idx = ent & ((1u << MDSF_ENTRY_BITS) - 1u);
        const int sx = (int)(ent >> MDSF_ENTRY_BITS) - 1;
        const AtomRec rec = recs[idx];
        const int f = (int)(idx / (unsigned)gp.natoms);
        const int Ax = tt.halfw[rec.type * 3], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
        int xlo, xhi;
        stamp_segment(rec.ir[0], Ax, nx, sx, xlo, xhi);
        // rows of this image inside the slab: destination planes [X0, X0+X)
        const int c0 = max(xlo - sx * nx, X0), c1 = min(xhi - sx * nx, min(X0 + sp.X, nx));
        const int nrows = max(c1 - c0, 0);
        const int i_first = c0 + sx * nx - (rec.ir[0] - Ax);
        sm.a[threadIdx.x] = make_int4((int)rec.tbase, rec.type | ((sx + 1) << 16) | ((f & 1) << 18), f >> 1,
                                      nrows | (i_first << 10) | ((c0 - X0) << 20));
        sm.b[threadIdx.x] = make_int4(rec.ir[1] - Ay, rec.ir[2] - Az, 2 * Ay, 2 * Az);
        if (!gp.separable) { sm.r[threadIdx.x * 3] = rec.r[0]; sm.r[threadIdx.x * 3 + 1] = rec.r[1]; sm.r[threadIdx.x * 3 + 2] = rec.r[2]; }
        sm.off[threadIdx.x + 1] = nrows * 2 * Ay;
    }
    if (threadIdx.x == 0) sm.off[0] = 0;
    __syncthreads();
    {   // block-wide inclusive scan of the <= 256 column counts (one per thread)
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int v = (int)threadIdx.x < ne ? sm.off[threadIdx.x + 1] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t0 = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t0; }
        if (lane == 31) sm.wsum[wid] = v;
        __syncthreads();
        int basev = 0;
        for (int w2 = 0; w2 < wid; ++w2) basev += sm.wsum[w2];
        if ((int)threadIdx.x < ne) sm.off[threadIdx.x + 1] = v + basev;
    }
    __syncthreads();
    const int total = sm.off[ne];
    for (int v = threadIdx.x; v < total; v += blockDim.x) {
        int lo = 0, hi = ne - 1;          // entry e with off[e] <= v < off[e+1]
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sm.off[mid] <= v) lo = mid; else hi = mid - 1; }
        const int4 a = sm.a[lo], b = sm.b[lo];
        const int twoAy = b.z, nzr = b.w;
        const int local = v - sm.off[lo];
        const int r = local / twoAy, j = local - r * twoAy;
        const int type = a.y & 0xffff, sx = ((a.y >> 16) & 3) - 1, part = (a.y >> 18) & 1;
        const int i = ((a.w >> 10) & 1023) + r, xl = (a.w >> 20) + r;
        // y image of this column
        const int py = b.x + j;
        const int sy = py < 0 ? -1 : (py >= ny ? 1 : 0);
        const int cy = py - sy * ny;
        const bool corner = (sx != 0 && sy != 0 && gp.fold_mode == 0);
        const int shlo = (corner && sy != -1) ? gp.nb : nz, shhi = (corner && sy != 1) ? -gp.nb : -nz;
        unsigned long long* colp = acc + ((((size_t)a.z * sp.X + xl) * ny + cy) * (size_t)nz) * 2 + part;
        const int pz0 = b.y;
        if (gp.separable) {
            const double* T = atom_tables + (unsigned)a.x;
            const int twoAx = 2 * tt.halfw[type * 3];
            double exy = T[i] * T[twoAx + j];
            if (tt.ctab != nullptr) exy *= tt.ctab[tt.ctab_off[type] + i * twoAy + j];
            exy *= sp.scale;
            const double* ez = T + twoAx + twoAy;
            for (int k = 0; k < nzr; ++k) {
                const int pz = pz0 + k;
                const int cz = pz < 0 ? pz + shlo : (pz >= nz ? pz + shhi : pz);
                const long long q = __double2ll_rn(exy * ez[k]);
                atomicAdd(colp + 2 * (size_t)cz, (unsigned long long)q);
            }
        } else {
            // general ucell: one exp per cell, exactly the reference's expression (dens.py:299-308)
            const double rx = sm.r[lo * 3], ry = sm.r[lo * 3 + 1], rz = sm.r[lo * 3 + 2];
            const int Ax = tt.halfw[type * 3];
            const int px = (int)(rx / gp.dr[0]) - Ax + i;
            const double bx = __dsub_rn(rx, __dmul_rn((double)px, gp.dr[0]));
            const double by = __dsub_rn(ry, __dmul_rn((double)py, gp.dr[1]));
            const double t2 = tt.two_sig2[type], amp = tt.amp[type] * sp.scale;
            for (int k = 0; k < nzr; ++k) {
                const int pz = pz0 + k;
                const int cz = pz < 0 ? pz + shlo : (pz >= nz ? pz + shhi : pz);
                const double bzv = __dsub_rn(rz, __dmul_rn((double)pz, gp.dr[2]));
                const double c0 = gp.u[0] * bx + gp.u[3] * by + gp.u[6] * bzv;
                const double c1 = gp.u[1] * bx + gp.u[4] * by + gp.u[7] * bzv;
                const double c2 = gp.u[2] * bx + gp.u[5] * by + gp.u[8] * bzv;
                const long long q = __double2ll_rn(amp * exp(-(c0 * c0 + c1 * c1 + c2 * c2) / t2));
                atomicAdd(colp + 2 * (size_t)cz, (unsigned long long)q);
            }
        }
    }
    // reductions are fire-and-forget: every thread must see its own performed before the grid-wide
    // barrier lets the z pass read the accumulator (grid.sync() only fences from one thread per block)
    __threadfence();
    __syncthreads();
}

// ---- K3z: one group of `ncol` columns: slab accumulator -> fp64 -> z FFT -> complex pair volume ---
// (clears the accumulator cells it reads; ncol consecutive (x,y) columns are contiguous in memory)
__device__ __forceinline__ void zpass_item(double* sre, double* sim, const double* twr, const double* twi,
                                           longlong2* __restrict__ acc, double2* __restrict__ vol, double2* __restrict__ dens_dump,
                                           const FftPlan& plan, const GridParams& gp, const SlabParams& sp, int slab, int q,
                                           long long col0, int ncol, int* __restrict__ err_flag)
{
    const int nz = gp.n[2], nzp = gp.nzp, pad = gp.pad_shift;
    const int X0 = slab * sp.X;
    const int xcount = min(sp.X, gp.n[0] - X0);
    const long long slab_cols = (long long)xcount * gp.n[1];
    const int nc = (int)min((long long)ncol, slab_cols - col0);
    longlong2* src = acc + ((long long)q * sp.X * gp.n[1] + col0) * nz;
    double2* dst = vol + (((long long)q * gp.n[0] + X0) * gp.n[1] + col0) * nz;
    long long overflow = 0;
    const int total = nc * nz;
    const int lnz = (nz & (nz - 1)) ? -1 : __ffs(nz) - 1;        // power-of-two nz: shifts instead of divides
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * blockDim.x) {      // 8 loads in flight per thread
        longlong2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < total ? __ldcg(src + i) : make_longlong2(0, 0); }   // L2 only: L1 is not coherent with the reductions
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) {
                __stcg(src + i, make_longlong2(0, 0));
                overflow |= (v[u].x ^ (v[u].x << 1)) | (v[u].y ^ (v[u].y << 1));     // bit 63 != bit 62: |value| >= 2^62
                const int c = lnz >= 0 ? (i >> lnz) : i / nz, z = i - c * nz;
                const int a = c * nzp + z + (z >> pad);
                sre[a] = (double)v[u].x * sp.inv_scale; sim[a] = (double)v[u].y * sp.inv_scale;
            }
        }
    }
    if (overflow < 0) atomicExch(err_flag, 2);
    __syncthreads();
    if (dens_dump != nullptr) {
        double2* dd = dens_dump + (((long long)q * gp.n[0] + X0) * gp.n[1] + col0) * nz;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int c = i / nz, z = i - c * nz;
            const int a = c * nzp + z + (z >> pad);
            dd[i] = make_double2(sre[a], sim[a]);
        }
        __syncthreads();     // the FFT below rewrites the tile in place
    }
    fft_tile_z(sre, sim, twr, twi, plan, nc, nzp, pad);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i / nz, z = i - c * nz;
        const int a = c * nzp + z + (z >> pad);
        dst[i] = make_double2(sre[a], sim[a]);
    }
    __threadfence();        // the cleared accumulator cells must be visible before the slab is scattered into again
    __syncthreads();
}

// ---- K3z (fast path, Nz = R*R): two radix-R stages, accumulator -> registers -> shared -> registers -> volume
// Stage 1 reads the fixed-point cells straight into registers (coalesced: lanes = consecutive z), clears
// them, converts, transforms and parks the result in shared memory; stage 2 transforms again and stores
// output k2 of block k1 at z = R*k2 + k1.  Transposing the digit-reversed result of two equal radices
// gives NATURAL frequency order, so the z axis needs no permutation table on this path.
template <int R>
__device__ __forceinline__ void zpass_item_fast(double* sre, double* sim, const double* twr, const double* twi,
                                                longlong2* __restrict__ acc, double2* __restrict__ vol, double2* __restrict__ dens_dump,
                                                const GridParams& gp, const SlabParams& sp, int slab, int q,
                                                long long col0, int ncol, int* __restrict__ err_flag)
{
    constexpr int NZ = R * R;
    const int nzp = gp.nzp, pad = gp.pad_shift;
    const int X0 = slab * sp.X;
    const int xcount = min(sp.X, gp.n[0] - X0);
    const long long slab_cols = (long long)xcount * gp.n[1];
    const int nc = (int)min((long long)ncol, slab_cols - col0);
    longlong2* src = acc + ((long long)q * sp.X * gp.n[1] + col0) * NZ;
    const long long vbase = (((long long)q * gp.n[0] + X0) * gp.n[1] + col0) * NZ;
    double2* dst = vol + vbase;
    long long overflow = 0;
    for (int it = threadIdx.x; it < nc * R; it += blockDim.x) {
        const int f = it / R, n2 = it - f * R;
        longlong2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = __ldcg(src + f * NZ + n2 + R * j);      // L2 only: L1 is not coherent with the reductions
        double xr[R], xi[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            __stcg(src + f * NZ + n2 + R * j, make_longlong2(0, 0));
            overflow |= (v[j].x ^ (v[j].x << 1)) | (v[j].y ^ (v[j].y << 1));
            xr[j] = (double)v[j].x * sp.inv_scale; xi[j] = (double)v[j].y * sp.inv_scale;
        }
        if (dens_dump != nullptr) {
#pragma unroll
            for (int j = 0; j < R; ++j) dens_dump[vbase + f * NZ + n2 + R * j] = make_double2(xr[j], xi[j]);
        }
        Dft<R>::run(xr, xi, twr, twi, NZ);
#pragma unroll
        for (int k = 1; k < R; ++k) {
            const double wr = twr[n2 * k], wi = twi[n2 * k];
            const double yr = xr[k] * wr - xi[k] * wi;
            xi[k] = xr[k] * wi + xi[k] * wr;
            xr[k] = yr;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int p = k * R + n2;
            const int a = f * nzp + p + (p >> pad);
            sre[a] = xr[k]; sim[a] = xi[k];
        }
    }
    if (overflow < 0) atomicExch(err_flag, 2);
    __syncthreads();
    for (int it = threadIdx.x; it < nc * R; it += blockDim.x) {
        const int f = it / R, b = it - f * R;
        double xr[R], xi[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int p = b * R + j;
            const int a = f * nzp + p + (p >> pad);
            xr[j] = sre[a]; xi[j] = sim[a];
        }
        Dft<R>::run(xr, xi, twr, twi, NZ);
#pragma unroll
        for (int k = 0; k < R; ++k) dst[f * NZ + R * k + b] = make_double2(xr[k], xi[k]);
    }
    __threadfence();
    __syncthreads();
}

// ---- persistent slab pipeline: static work queue + per-slab completion counters, no grid-wide barriers
// The queue is laid out by scan_slabs_kernel: step t = scatter items of slab t, then z-pass items of slab
// t-1.  Items are dealt round-robin (item = k*gridDim + blockIdx).  A z-pass item waits until every
// scatter item of its slab has signalled; a scatter item of slab s waits until the z pass has drained
// slab s-R out of its ring buffer (R slab accumulators stay resident in L2).  Every CTA walks its items in
// increasing order and only ever waits for EARLIER items, and all CTAs are co-resident (cooperative
// launch), so the smallest unfinished item is never blocked: no deadlock.  Spins are bounded and abort
// the whole launch through ctl[1].  The number of items in flight (= CTAs) must stay below ring x (items
// per slab), otherwise most CTAs spin: 2 CTAs of 256 threads per SM measured best on B200.
#ifndef MDSF_PIPE_MINBLOCKS
#define MDSF_PIPE_MINBLOCKS 2
#endif
#ifndef MDSF_PIPE_THREADS
#define MDSF_PIPE_THREADS 256
#endif
#define MDSF_PIPE_SPIN_LIMIT (1u << 24)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// thread 0 spins until *ctr >= target (or the launch is aborted); returns false on abort
__device__ __forceinline__ bool wait_counter(const unsigned* ctr, unsigned target, unsigned* ctl) {
    __shared__ int ok;
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        int good = 1;
        while (ld_acquire_u32(ctr) < target) {
            __nanosleep(64);
            if (++spins > MDSF_PIPE_SPIN_LIMIT || ld_acquire_u32(ctl + 1) != 0) { atomicExch(ctl + 1, 1u); good = 0; break; }
        }
        ok = good;
    }
    __syncthreads();
    const bool r = ok != 0;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(MDSF_PIPE_THREADS, MDSF_PIPE_MINBLOCKS)
slab_pipeline_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ entries,
                     const unsigned* __restrict__ slab_start, const unsigned* __restrict__ step_start,
                     unsigned* __restrict__ ctl,          // [0] unused, [1] abort, [2..2+n) scatter done, [2+n..2+2n) z-pass done
                     const double* __restrict__ atom_tables, unsigned long long* __restrict__ acc, size_t acc_cells, int ring,
                     double2* __restrict__ vol, double2* __restrict__ dens_dump, FftPlan zplan, const double2* __restrict__ twz,
                     GridParams gp, TypeTable tt, SlabParams sp, int npairs, int ncol, int zfast, int* __restrict__ err_flag)
{
    extern __shared__ double smem[];
    __shared__ ScatterSmem ssm;
    const int nz = gp.n[2];
    double* sre = smem;
    double* sim = sre + (size_t)ncol * gp.nzp;
    double* twr = sim + (size_t)ncol * gp.nzp;
    double* twi = twr + nz;
    if (zplan.nstages > 0) load_twiddles(twr, twi, twz, nz);
    __syncthreads();
    unsigned* scat_done = ctl + 2;
    unsigned* zp_done = ctl + 2 + sp.nslabs;
    const unsigned total_items = step_start[sp.nslabs + 1];
    int t = 0;
    for (unsigned item = blockIdx.x; item < total_items; item += gridDim.x) {
        while (item >= step_start[t + 1]) ++t;
        const unsigned local = item - step_start[t];
        unsigned beg = 0, end = 0;
        if (t < sp.nslabs) { beg = slab_start[t]; end = slab_start[t + 1]; }
        const unsigned nA = (end - beg + MDSF_SC_ENTRIES - 1) / MDSF_SC_ENTRIES;
        if (local < nA) {
            // ---- scatter item `local` of slab t into ring buffer t % ring
            if (t >= ring) {
                const int sd = t - ring;
                const int xc = min(sp.X, gp.n[0] - sd * sp.X);
                const unsigned need = (unsigned)(((long long)xc * gp.n[1] + ncol - 1) / ncol) * npairs;
                if (!wait_counter(zp_done + sd, need, ctl)) break;
            }
            const unsigned cb = beg + local * MDSF_SC_ENTRIES;
            if (!(gp.debug_skip & 32))
                scatter_chunk(ssm, recs, entries, cb, (int)min((unsigned)MDSF_SC_ENTRIES, end - cb), atom_tables,
                              acc + (size_t)(t % ring) * acc_cells * 2, gp, tt, sp, t);
            if (threadIdx.x == 0) atomicAdd(scat_done + t, 1u);
        } else {
            // ---- z-pass item of slab t-1
            const int sd = t - 1;
            const unsigned need = (slab_start[sd + 1] - slab_start[sd] + MDSF_SC_ENTRIES - 1) / MDSF_SC_ENTRIES;
            if (!wait_counter(scat_done + sd, need, ctl)) break;
            const int xc = min(sp.X, gp.n[0] - sd * sp.X);
            const int groups = (int)(((long long)xc * gp.n[1] + ncol - 1) / ncol);
            const int zi = (int)(local - nA);
            const int q = zi / groups, g = zi - q * groups;
            longlong2* abuf = reinterpret_cast<longlong2*>(acc + (size_t)(sd % ring) * acc_cells * 2);
            if (!(gp.debug_skip & 64)) {
                if (zfast == 16) zpass_item_fast<16>(sre, sim, twr, twi, abuf, vol, dens_dump, gp, sp, sd, q, (long long)g * ncol, ncol, err_flag);
                else if (zfast == 8) zpass_item_fast<8>(sre, sim, twr, twi, abuf, vol, dens_dump, gp, sp, sd, q, (long long)g * ncol, ncol, err_flag);
                else zpass_item(sre, sim, twr, twi, abuf, vol, dens_dump, zplan, gp, sp, sd, q, (long long)g * ncol, ncol, err_flag);
            }
            if (threadIdx.x == 0) atomicAdd(zp_done + sd, 1u);
        }
    }
    if (threadIdx.x == 0 && ld_acquire_u32(ctl + 1) != 0) atomicExch(err_flag, 3);
}
