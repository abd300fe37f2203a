// K1/K2: per-atom preparation and deterministic binning of (atom image, tile) pairs.
#pragma once
#include "mdsf_common.cuh"

// ---- K1: rescale (dens.py:56-58) o wrap (dens.py:209-221) o cell index (dens.py:285) ----------
// C = dtype of the coordinate array, P = dtype numpy promotes (coords, dims) to.  Every
// arithmetic step is done in P and rounded back to C exactly where numpy stores into the
// coords array, so the wrapped coordinates and the cell indices are bit-identical.
template <typename P> __device__ __forceinline__ P mul_rn(P a, P b);
template <> __device__ __forceinline__ float  mul_rn<float>(float a, float b)   { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename P> __device__ __forceinline__ P add_rn(P a, P b);
template <> __device__ __forceinline__ float  add_rn<float>(float a, float b)   { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// (image, tile, slab) pairs of one atom: for every periodic image (sx,sy) of the stamp's xy
// rectangle, every tile it overlaps, every z slab it touches.  emit_pairs_kernel walks the same loops.
__device__ __forceinline__ unsigned count_pairs(const AtomRec& rec, const GridParams& gp, const TypeTable& tt) {
    const int Ax = tt.halfw[rec.type * 3], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
    unsigned total = 0;
    for (int sx = -1; sx <= 1; ++sx) {
        int xlo, xhi;
        stamp_segment(rec.ir[0], Ax, gp.n[0], sx, xlo, xhi);
        if (xhi <= xlo) continue;
        const int ntx = (xhi - 1 - sx * gp.n[0]) / gp.tx - (xlo - sx * gp.n[0]) / gp.tx + 1;
        for (int sy = -1; sy <= 1; ++sy) {
            int ylo, yhi;
            stamp_segment(rec.ir[1], Ay, gp.n[1], sy, ylo, yhi);
            if (yhi <= ylo) continue;
            const int nty = (yhi - 1 - sy * gp.n[1]) / gp.ty - (ylo - sy * gp.n[1]) / gp.ty + 1;
            int shlo, shhi, kA, kB;
            const unsigned sm = image_slabmask(rec.ir[2], Az, sx, sy, gp.n[2], gp.nb, gp.fold_mode, gp.zs, shlo, shhi, kA, kB);
            total += (unsigned)(ntx * nty * __popc(sm));
        }
    }
    return total;
}

template <typename F>
__device__ __forceinline__ void for_each_pair(const AtomRec& rec, int a, int f, const GridParams& gp, const TypeTable& tt, F&& fn) {
    const int Ax = tt.halfw[rec.type * 3 + 0], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
    const int ntiles = gp.ntx * gp.nty;
    const int ltx = __ffs(gp.tx) - 1, lty = __ffs(gp.ty) - 1;     // tile sizes are powers of two; destinations are >= 0
    for (int sx = -1; sx <= 1; ++sx) {
        int xlo, xhi;
        stamp_segment(rec.ir[0], Ax, gp.n[0], sx, xlo, xhi);
        if (xhi <= xlo) continue;
        const int tx0 = (xlo - sx * gp.n[0]) >> ltx, tx1 = (xhi - 1 - sx * gp.n[0]) >> ltx;
        for (int sy = -1; sy <= 1; ++sy) {
            int ylo, yhi;
            stamp_segment(rec.ir[1], Ay, gp.n[1], sy, ylo, yhi);
            if (yhi <= ylo) continue;
            const int ty0 = (ylo - sy * gp.n[1]) >> lty, ty1 = (yhi - 1 - sy * gp.n[1]) >> lty;
            const unsigned payload = (unsigned)a | ((unsigned)(sx + 1) << MDSF_ATOM_BITS) |
                                     ((unsigned)(sy + 1) << (MDSF_ATOM_BITS + 2));
            unsigned sm;
            if (gp.nslab == 1) {
                sm = Az > 0 ? 1u : 0u;          // one list per tile (tile mode): every image with cells belongs to it
            } else {
                int shlo, shhi, kA, kB;
                sm = image_slabmask(rec.ir[2], Az, sx, sy, gp.n[2], gp.nb, gp.fold_mode, gp.zs, shlo, shhi, kA, kB);
            }
            for (int tX = tx0; tX <= tx1; ++tX)
                for (int tY = ty0; tY <= ty1; ++tY) {
                    const unsigned kbase = (unsigned)(f * ntiles + tX * gp.nty + tY) * (unsigned)gp.nslab;
                    for (unsigned m = sm; m; m &= m - 1) fn(kbase + (unsigned)(__ffs(m) - 1), payload);
                }
        }
    }
}

#define MDSF_PREP_STAGE 512        // doubles of factor tables one warp stages in shared memory (c2: 32 atoms x 12)

template <typename C, typename P>
__global__ void __launch_bounds__(256)
prep_atoms_kernel(C* __restrict__ coords,            // [nframes][natoms][3], rewritten in place
                  const int* __restrict__ type_id,   // [natoms]
                  AtomRec* __restrict__ recs,        // [nframes][natoms]
                  unsigned* __restrict__ pair_count, // [nframes*natoms]
                  double* __restrict__ tables,       // [nframes][tstride] per-atom Gaussian factor tables
                  GridParams gp, TypeTable tt, BatchScales sc, int nframes,
                  long long wrap_lo, long long wrap_hi, int* __restrict__ err_flag,
                  unsigned* __restrict__ tile_counter /* direct binning: list lengths per (frame, tile) key, or nullptr */,
                  int mono, double mono_sin, double mono_cos /* monoclinic pre-transform (main_gromacs.py:204-207) */)
{
    // A lane's record (48 B) and factor tables (16 (Ax+Ay+Az) B) are contiguous with its neighbours' in global memory
    // but strided across the lanes of a store instruction; both are staged per warp in shared memory and written
    // out as whole 256-byte rows (the strided form touched every 32-byte sector four times).
    __shared__ double s_tab[8][MDSF_PREP_STAGE];
    __shared__ double s_rec[8][32 * 6];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)nframes * gp.natoms;
    const long long nloop = (total + 31) / 32 * 32;                       // whole warps stay in the loop together
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nloop;
         idx += (long long)gridDim.x * blockDim.x) {
        const bool live = idx < total;
        const int f = live ? (int)(idx / gp.natoms) : 0;
        const int a = live ? (int)(idx - (long long)f * gp.natoms) : 0;
        const int t = live ? type_id[a] : 0;
        AtomRec rec{};
        rec.type = t;
        rec.tbase = live ? (unsigned)((long long)f * gp.tstride + tt.toff[a]) : 0u;
        rec.pad_ = 0;
        bool bad = false;
        if (live) {
            C* src = coords + idx * 3;
            C v[3] = {src[0], src[1], src[2]};
            if (mono) {
                // T[..., 1] = T[..., 1] / np.sin(theta); T[..., 0] = T[..., 0] - T[..., 1] * np.cos(theta)
                // (main_gromacs.py:206-207): np.sin / np.cos return float64 scalars, so numpy evaluates both lines
                // in float64 and rounds to the coords dtype on assignment; no fused multiply-add
                v[1] = (C)((double)v[1] / mono_sin);
                v[0] = (C)__dsub_rn((double)v[0], __dmul_rn((double)v[1], mono_cos));
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                // rc[it,:,i] *= a[it,i]   (dens.py:58)
                C r = (C)mul_rn<P>((P)v[d], (P)sc.a[f][d]);
                if (a >= wrap_lo && a < wrap_hi) {
                    const P L = (P)gp.box[d];
                    // np.where(r < L, r, r - L) then np.where(r > 0, r, r + L)   (dens.py:211-212)
                    if (!((P)r < L)) r = (C)add_rn<P>((P)r, -L);
                    if (!((P)r > (P)0)) r = (C)add_rn<P>((P)r, L);
                }
                src[d] = r;
                const double rd = (double)r;
                const double q = rd / gp.dr[d];            // IEEE fp64 divide, as numpy (dens.py:285)
                const int A = tt.halfw[t * 3 + d];
                int ir = 0;
                if (!(q > -2147483000.0 && q < 2147483000.0)) bad = true;   // also catches NaN
                else ir = (int)q;                          // astype(int): truncation toward zero
                // the stamp [ir-A, ir+A) must stay inside the padded grid [-B, N+B)  (dens.py:292-297)
                if (ir - A < -gp.nb || ir + A > gp.n[d] + gp.nb) bad = true;
                rec.r[d] = rd;
                rec.ir[d] = ir;
            }
        }
        {   // records: 6 eight-byte words per lane -> 192 contiguous words per warp
            const double* w = reinterpret_cast<const double*>(&rec);
#pragma unroll
            for (int i = 0; i < 6; ++i) s_rec[warp][lane * 6 + i] = w[i];
            __syncwarp();
            const long long idx0 = idx - lane;
            const int nvalid = (int)min(32LL, total - idx0) * 6;
            double* dst = reinterpret_cast<double*>(recs + idx0);
            for (int j = lane; j < nvalid; j += 32) dst[j] = s_rec[warp][j];
            __syncwarp();
        }
        unsigned npairs = 0;
        if (live && bad) atomicExch(err_flag, 1);
        const bool ok = live && !bad;
        if (ok) {
            if (tile_counter != nullptr) for_each_pair(rec, a, f, gp, tt, [&](unsigned key, unsigned) { atomicAdd(tile_counter + key, 1u); ++npairs; });
            else npairs = count_pairs(rec, gp, tt);
        }
        if (live) pair_count[idx] = npairs;
        if (!gp.separable) continue;                       // warp-uniform
        // One-dimensional Gaussian factors of this atom's stamp (dens.py:299-308 factorised):
        //   exp(-|c|^2/(2s^2)) = EX[i] * EY[j] * C_type[i][j] * EZ[k]/amp, with
        //   EX[i] = exp(-(cxx bx_i^2 + 2 gxy bx_i by_0)/(2s^2)),  EY[j] = exp(-(cyy by_j^2 - 2 gxy (j dy) bx_0)/(2s^2)),
        //   EZ[k] = Nel/s^3 exp(-czz bz_k^2/(2s^2)),  C[i][j] = exp(-2 gxy dx dy i j/(2s^2)) (per type, host-built).
        // b = r - (i - B)*dr with the product rounded on its own, as numpy does (dens.py:252-256,299).
        const int Ax = tt.halfw[t * 3], Ay = tt.halfw[t * 3 + 1], Az = tt.halfw[t * 3 + 2];
        const unsigned size = ok ? (unsigned)(2 * (Ax + Ay + Az)) : 0u;
        const unsigned base = __reduce_min_sync(0xffffffffu, ok ? rec.tbase : 0xffffffffu);
        const unsigned off = ok ? rec.tbase - base : 0u;
        const unsigned span = __reduce_max_sync(0xffffffffu, off + size);
        const bool staged = span <= MDSF_PREP_STAGE;       // false where a warp straddles two frames or holds heavy ions
        if (ok) {
            // one reciprocal per atom instead of a divide per table entry (the argument moves by <= 1 ulp: 1e-16 of a
            // term; the golden-vector density tolerance is 1e-13 of the peak)
            const double it2 = 1.0 / tt.two_sig2[t], amp = tt.amp[t];
            double* T = staged ? &s_tab[warp][off] : tables + rec.tbase;
            const double bx0 = __dsub_rn(rec.r[0], __dmul_rn((double)(rec.ir[0] - Ax), gp.dr[0]));
            const double by0 = __dsub_rn(rec.r[1], __dmul_rn((double)(rec.ir[1] - Ay), gp.dr[1]));
            for (int i = 0; i < 2 * Ax; ++i) {
                const double b = __dsub_rn(rec.r[0], __dmul_rn((double)(rec.ir[0] - Ax + i), gp.dr[0]));
                T[i] = exp(-(gp.cxx * b * b + 2.0 * gp.gxy * b * by0) * it2);
            }
            T += 2 * Ax;
            for (int j = 0; j < 2 * Ay; ++j) {
                const double b = __dsub_rn(rec.r[1], __dmul_rn((double)(rec.ir[1] - Ay + j), gp.dr[1]));
                T[j] = exp(-(gp.cyy * b * b - 2.0 * gp.gxy * ((double)j * gp.dr[1]) * bx0) * it2);
            }
            T += 2 * Ay;
            for (int k = 0; k < 2 * Az; ++k) {
                const double b = __dsub_rn(rec.r[2], __dmul_rn((double)(rec.ir[2] - Az + k), gp.dr[2]));
                T[k] = amp * exp(-(gp.czz * b * b) * it2);
            }
        }
        __syncwarp();
        if (staged && span > 0) {
            // tables of atoms the stamp check rejected (holes inside the span) receive stale staging data; nobody reads them
            double* dst = tables + base;
            for (unsigned j = lane; j < span; j += 32) dst[j] = s_tab[warp][j];
        }
        __syncwarp();
    }
}

// ---- K2a: emit (key = (frame*ntiles + tile)*nslab + slab, payload = atom | sx | sy) at scanned offsets
__global__ void __launch_bounds__(256)
emit_pairs_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ pair_count,
                  const unsigned* __restrict__ pair_off, unsigned* __restrict__ keys,
                  unsigned* __restrict__ vals, GridParams gp, TypeTable tt, int nframes)
{
    const long long total = (long long)nframes * gp.natoms;
    const int ntiles = gp.ntx * gp.nty;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        if (pair_count[idx] == 0) continue;
        const int f = (int)(idx / gp.natoms);
        const int a = (int)(idx - (long long)f * gp.natoms);
        const AtomRec rec = recs[idx];
        const int Ax = tt.halfw[rec.type * 3 + 0], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
        unsigned o = pair_off[idx];
        for (int sx = -1; sx <= 1; ++sx) {
            int xlo, xhi;
            stamp_segment(rec.ir[0], Ax, gp.n[0], sx, xlo, xhi);
            if (xhi <= xlo) continue;
            const int tx0 = (xlo - sx * gp.n[0]) / gp.tx, tx1 = (xhi - 1 - sx * gp.n[0]) / gp.tx;
            for (int sy = -1; sy <= 1; ++sy) {
                int ylo, yhi;
                stamp_segment(rec.ir[1], Ay, gp.n[1], sy, ylo, yhi);
                if (yhi <= ylo) continue;
                const int ty0 = (ylo - sy * gp.n[1]) / gp.ty, ty1 = (yhi - 1 - sy * gp.n[1]) / gp.ty;
                const unsigned payload = (unsigned)a | ((unsigned)(sx + 1) << MDSF_ATOM_BITS) |
                                         ((unsigned)(sy + 1) << (MDSF_ATOM_BITS + 2));
                int shlo, shhi, kA, kB;
                const unsigned sm = image_slabmask(rec.ir[2], Az, sx, sy, gp.n[2], gp.nb, gp.fold_mode, gp.zs, shlo, shhi, kA, kB);
                for (int tX = tx0; tX <= tx1; ++tX)
                    for (int tY = ty0; tY <= ty1; ++tY) {
                        const unsigned kbase = (unsigned)(f * ntiles + tX * gp.nty + tY) * (unsigned)gp.nslab;
                        for (unsigned m = sm; m; m &= m - 1) {
                            keys[o] = kbase + (unsigned)(__ffs(m) - 1);
                            vals[o] = payload;
                            ++o;
                        }
                    }
            }
        }
    }
}

// ---- K2 (tile splat mode): direct binning.  The fixed-point tile accumulation is order-independent, so the
// lists need no stable sort: count the pairs of every (frame, tile) key with integer atomics, scan the counts
// into list starts, and let every pair claim a slot of its list from a per-key cursor.  Replaces scan + fill +
// emit + radix sort + list starts (0.43 ms -> 0.15 ms per 32 c2 frames, serialised ncu times).
// PLACE = false: list lengths; PLACE = true: payloads into their lists (start = scanned lengths, cursor zeroed)
template <bool PLACE>
__global__ void __launch_bounds__(256)
bin_pairs_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ pair_count, unsigned* __restrict__ counter,
                 const unsigned* __restrict__ start, unsigned* __restrict__ vals, uint4* __restrict__ prec,
                 GridParams gp, TypeTable tt, int nframes)
{
    const long long total = (long long)nframes * gp.natoms;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        if (pair_count[idx] == 0) continue;
        const int f = (int)(idx / gp.natoms);
        const int a = (int)(idx - (long long)f * gp.natoms);
        const AtomRec rec = recs[idx];
        for_each_pair(rec, a, f, gp, tt, [&](unsigned key, unsigned payload) {
            if (PLACE) {
                const unsigned pos = start[key] + atomicAdd(counter + key, 1u);
                if (prec != nullptr) {       // 16-byte pair record (layout: splat_zfft_kernel, PREC)
                    prec[pos] = make_uint4((unsigned)(rec.ir[0] + 1024) | ((unsigned)(rec.ir[1] + 1024) << 12) | ((payload >> MDSF_ATOM_BITS) << 24),
                                           (unsigned)(rec.ir[2] + 1024) | ((unsigned)rec.type << 13), rec.tbase, (unsigned)a);
                } else {
                    vals[pos] = payload;
                }
            } else {
                atomicAdd(counter + key, 1u);
            }
        });
    }
}

__global__ void fill_u32_kernel(unsigned* __restrict__ p, unsigned v, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// ---- K2c: list boundaries in the sorted key array: start[k] = first index with key >= k -----
__global__ void __launch_bounds__(256)
tile_starts_kernel(const unsigned* __restrict__ keys, long long n, unsigned nkeys,
                   unsigned* __restrict__ start /* [nkeys+1] */)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i <= n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long kprev = (i == 0) ? -1 : (long long)min(keys[i - 1], nkeys);
        const long long kcur = (i == n) ? (long long)nkeys : (long long)min(keys[i], nkeys);
        for (long long k = kprev + 1; k <= kcur; ++k) start[k] = (unsigned)i;
    }
}
