// K1/K2: per-atom preparation and binning of (atom image, tile, z slab) pairs.
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

// Per-atom factor tables (one block per atom and frame, K1 writes them):
//   [TX-1 zeros][EX: 2Ax][TX-1 zeros] [TY-1 zeros][EY: 2Ay][TY-1 zeros] [EZ: 2Az]
// The zero pads let a splat warp read EX / EY for EVERY column of its tile without a clip test: columns outside the
// atom's stamp fall into a pad and contribute exactly zero.
__host__ __device__ inline int table_doubles(int lcol, int Ax, int Ay, int Az) {
    const int TX = 1 << ((lcol + 1) >> 1), TY = 1 << (lcol >> 1);
    return 2 * (Ax + Ay + Az) + 2 * (TX - 1) + 2 * (TY - 1);
}

// The k-th tile (width 2^lt) the folded stamp [ir-A, ir+A) touches along one dimension, counted over its fold images
// in the order side = -1, 0, +1 (low padding / cell / high padding, dens.py:86-108).  Returns false when the stamp
// touches fewer tiles.  d0, d1: destination range of that image; ist: stamp index of d0.
__device__ __forceinline__ bool axis_slot(int ir, int A, int N, int lt, int k, int& side, int& tile, int& d0, int& d1, int& ist) {
#pragma unroll
    for (int s = -1; s <= 1; ++s) {
        int plo, phi;
        stamp_segment(ir, A, N, s, plo, phi);
        if (phi > plo) {
            const int a0 = plo - s * N, a1 = phi - s * N;
            const int t0 = a0 >> lt, nt = ((a1 - 1) >> lt) - t0 + 1;
            if (k < nt) { side = s; tile = t0 + k; d0 = a0; d1 = a1; ist = plo - (ir - A); return true; }
            k -= nt;
        }
    }
    return false;
}

// Every (image, tile, slab) pair of one atom inside ONE (x tile, y tile) column of tiles.  The stamp on the padded grid
// is cut into its <= 27 fold images; each image is one box in destination space with ONE shift, so the splat never
// sees segments or the corner rule again.  fn(key, rec, aux): key = (frame * ntiles + tile) * nslab + slab.
template <typename F>
__device__ __forceinline__ void pairs_of_tile(const AtomRec& rec, int a, int f, const GridParams& gp, const TypeTable& tt,
                                              int Ax, int Ay, int Az, int sx, int tX, int dx0, int dx1, int ix0,
                                              int sy, int tY, int dy0, int dy1, int jy0, F&& fn) {
    const int ltx = (gp.lcol + 1) >> 1, lty = gp.lcol >> 1;
    const int TX = 1 << ltx, TY = 1 << lty;
    const int ZW = gp.zw, lzw = 31 - __clz(ZW);                  // slab width (a power of two)
    const int X0 = tX << ltx, Y0 = tY << lty;
    const int cx0 = max(dx0 - X0, 0), cy0 = max(dy0 - Y0, 0);
    const int i0 = ix0 + (X0 + cx0 - dx0), j0 = jy0 + (Y0 + cy0 - dy0);
    const unsigned kbase = (unsigned)(f * gp.ntx * gp.nty + tX * gp.nty + tY) * (unsigned)gp.nslab;
    const int ex0 = (int)rec.tbase + (TX - 1);                                   // index of EX[0]
    const int ey0 = (int)rec.tbase + 2 * (TX - 1) + 2 * Ax + (TY - 1);           // index of EY[0]
    const int ez0 = (int)rec.tbase + 2 * (TX - 1) + 2 * Ax + 2 * (TY - 1) + 2 * Ay;
    const int ctbase = (tt.ctab != nullptr) ? tt.ctab_off[rec.type] : 0;
    for (int sz = -1; sz <= 1; ++sz) {
        int zlo, zhi;
        stamp_segment(rec.ir[2], Az, gp.n[2], sz, zlo, zhi);
        if (zhi <= zlo) continue;
        const int shz = fold_shift_z(sx, sy, sz, gp.n[2], gp.nb, gp.fold_mode);
        const int dz0 = zlo + shz, dz1 = zhi + shz;
        const int kz0 = zlo - (rec.ir[2] - Az);
        for (int s = dz0 >> lzw; s <= (dz1 - 1) >> lzw; ++s) {
            const int Z0 = s << lzw;
            const int zoff = max(dz0 - Z0, 0), zend = min(dz1 - Z0, ZW);
            const int k0 = kz0 + (Z0 + zoff - dz0);
            PairRec pr;
            PairAux pa;
            if (gp.separable) {
                pr.x = (unsigned)(ez0 + k0 - zoff);
                pr.y = (unsigned)(ex0 + i0 - cx0);
                pr.z = (unsigned)(ey0 + j0 - cy0);
                pr.w = (unsigned)(zoff | ((zend - zoff) << 7));
                pa.x = (unsigned)(ctbase + (i0 - cx0) * 2 * Ay + (j0 - cy0));
                pa.y = (unsigned)(2 * Ay);
            } else {
                const int cx1 = min(dx1 - X0, TX), cy1 = min(dy1 - Y0, TY);
                pr.x = (unsigned)(Z0 - shz);                  // padded z of slab-local cell 0
                pr.y = (unsigned)(X0 + sx * gp.n[0]);         // padded x of tile column 0
                pr.z = (unsigned)(Y0 + sy * gp.n[1]);
                pr.w = (unsigned)(cx0 | (cx1 << 3) | (cy0 << 7) | (cy1 << 10) | (zoff << 14) | (zend << 21));
                pa.x = (unsigned)a;
                pa.y = 0u;
            }
            fn(kbase + (unsigned)s, pr, pa);
        }
    }
}

// number of (image, tile) slots axis_slot() enumerates along one dimension
__device__ __forceinline__ int axis_count(int ir, int A, int N, int lt) {
    int cnt = 0;
#pragma unroll
    for (int s = -1; s <= 1; ++s) {
        int plo, phi;
        stamp_segment(ir, A, N, s, plo, phi);
        if (phi > plo) cnt += ((phi - 1 - s * N) >> lt) - ((plo - s * N) >> lt) + 1;
    }
    return cnt;
}

// All pairs of 32 consecutive atoms of the walk order, spread evenly over the lanes of one warp.  One thread per atom
// left most lanes idle: a sodium ion (34^3 cells at c3) has ~500 pairs, a hydrogen ~10, and every fifth warp holds an
// ion.  Work unit = one (x tile, y tile) slot of one atom (its z slabs are a short loop); units are numbered through
// an inclusive scan of the per-atom slot counts and lane l takes units l, l + 32, ...; the owner of a unit is found by
// a 5-step search over the scanned counts (shuffles), its record sits in shared memory.
template <typename F>
__device__ __forceinline__ void walk_pairs_warp(const AtomRec* __restrict__ recs, const unsigned* __restrict__ perm,
                                                long long slot0, long long total, const GridParams& gp, const TypeTable& tt,
                                                AtomRec* __restrict__ s_rec /* [32] */, unsigned* __restrict__ s_idx /* [2][32]: atom index, frame */, F&& fn) {
    const int lane = threadIdx.x & 31;
    const int ltx = (gp.lcol + 1) >> 1, lty = gp.lcol >> 1;
    int nys = 1, nu = 0;
    __syncwarp();
    if (slot0 + lane < total) {
        const long long idx = perm != nullptr ? (long long)perm[slot0 + lane] : slot0 + lane;
        const uint4* src = reinterpret_cast<const uint4*>(recs + idx);
        uint4* dst = reinterpret_cast<uint4*>(s_rec + lane);
        dst[0] = __ldcs(src); dst[1] = __ldcs(src + 1); dst[2] = __ldcs(src + 2);
        const int f0 = (int)(idx / gp.natoms);                  // (once per atom, not once per work unit)
        s_idx[lane] = (unsigned)(idx - (long long)f0 * gp.natoms);
        s_idx[32 + lane] = (unsigned)f0;
        const AtomRec& rec = s_rec[lane];
        const int Ax = tt.halfw[rec.type * 3 + 0], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
        if (!rec.pad_ && Ax > 0 && Ay > 0 && Az > 0) {       // atoms K1 rejected have no pairs
            nys = axis_count(rec.ir[1], Ay, gp.n[1], lty);
            nu = axis_count(rec.ir[0], Ax, gp.n[0], ltx) * nys;
        }
    }
    int incl = nu;                                             // inclusive scan of the unit counts
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    for (int u0 = 0; u0 < tot; u0 += 32) {
        const int u = u0 + lane;
        int o = 0;                                             // owner: first lane with incl > u
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, o + d - 1);
            if (v <= u) o += d;
        }
        const int excl_o = __shfl_sync(0xffffffffu, incl - nu, o & 31);
        const int nys_o = __shfl_sync(0xffffffffu, nys, o & 31);
        if (u >= tot) continue;
        const AtomRec rec = s_rec[o];
        const int a = (int)s_idx[o], f = (int)s_idx[32 + o];
        const int Ax = tt.halfw[rec.type * 3 + 0], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
        const int local = u - excl_o, kx = local / nys_o, ky = local - kx * nys_o;
        int sx, tX, dx0, dx1, ix0, sy, tY, dy0, dy1, jy0;
        axis_slot(rec.ir[0], Ax, gp.n[0], ltx, kx, sx, tX, dx0, dx1, ix0);
        axis_slot(rec.ir[1], Ay, gp.n[1], lty, ky, sy, tY, dy0, dy1, jy0);
        pairs_of_tile(rec, a, f, gp, tt, Ax, Ay, Az, sx, tX, dx0, dx1, ix0, sy, tY, dy0, dy1, jy0, fn);
    }
}

// x bucket of an atom's home cell (= its x tile, clamped): K2 walks the atoms bucket by bucket
__device__ __forceinline__ int x_bucket(int ir0, const GridParams& gp) {
    return min(max(ir0, 0) >> ((gp.lcol + 1) >> 1), gp.ntx - 1);
}

#define MDSF_PREP_STAGE 512        // doubles of factor tables one warp stages in shared memory (c2: 32 atoms x 12)

// K1: records and factor tables (and the number of atoms per x bucket for the walk order of K2).
template <typename C, typename P>
__global__ void __launch_bounds__(256)
prep_atoms_kernel(C* __restrict__ coords,            // [nframes][natoms][3], rewritten in place
                  const int* __restrict__ type_id,   // [natoms]
                  AtomRec* __restrict__ recs,        // [nframes][natoms]
                  double* __restrict__ tables,       // [nframes][tstride] per-atom Gaussian factor tables
                  GridParams gp, TypeTable tt, BatchScales sc, int nframes,
                  long long wrap_lo, long long wrap_hi, int* __restrict__ err_flag,
                  unsigned* __restrict__ xcnt        /* [nframes][ntx] atoms per x bucket of their home cell (nullptr: K2 keeps the atom order) */,
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
        rec.pad_ = bad ? 1u : 0u;                          // bin_place_kernel skips rejected atoms
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
        if (live && bad) atomicExch(err_flag, 1);
        if (live && xcnt != nullptr) atomicAdd(xcnt + f * gp.ntx + x_bucket(rec.ir[0], gp), 1u);
        const bool ok = live && !bad;
        if (!gp.separable) continue;                       // warp-uniform
        // One-dimensional Gaussian factors of this atom's stamp (dens.py:299-308 factorised):
        //   exp(-|c|^2/(2s^2)) = EX[i] * EY[j] * C_type[i][j] * EZ[k]/amp, with
        //   EX[i] = exp(-(cxx bx_i^2 + 2 gxy bx_i by_0)/(2s^2)),  EY[j] = exp(-(cyy by_j^2 - 2 gxy (j dy) bx_0)/(2s^2)),
        //   EZ[k] = 2^(52-e) Nel/s^3 exp(-czz bz_k^2/(2s^2)),  C[i][j] = exp(-2 gxy dx dy i j/(2s^2)) (per type, host-built).
        // EZ carries the fixed-point scale of the splat accumulators (a power of two: exact).
        // b = r - (i - B)*dr with the product rounded on its own, as numpy does (dens.py:252-256,299).
        const int Ax = tt.halfw[t * 3], Ay = tt.halfw[t * 3 + 1], Az = tt.halfw[t * 3 + 2];
        const int padx = (1 << ((gp.lcol + 1) >> 1)) - 1, pady = (1 << (gp.lcol >> 1)) - 1;
        const unsigned size = ok ? (unsigned)table_doubles(gp.lcol, Ax, Ay, Az) : 0u;
        const unsigned base = __reduce_min_sync(0xffffffffu, ok ? rec.tbase : 0xffffffffu);
        const unsigned off = ok ? rec.tbase - base : 0u;
        const unsigned span = __reduce_max_sync(0xffffffffu, off + size);
        const bool staged = span <= MDSF_PREP_STAGE;       // false where a warp straddles two frames or holds heavy ions
        if (ok) {
            // one reciprocal per atom instead of a divide per table entry (the argument moves by <= 1 ulp: 1e-16 of a
            // term; the golden-vector density tolerance is 1e-13 of the peak)
            const double it2 = 1.0 / tt.two_sig2[t], amp = tt.amp[t] * gp.fx_scale;
            double* T = staged ? &s_tab[warp][off] : tables + rec.tbase;
            const double bx0 = __dsub_rn(rec.r[0], __dmul_rn((double)(rec.ir[0] - Ax), gp.dr[0]));
            const double by0 = __dsub_rn(rec.r[1], __dmul_rn((double)(rec.ir[1] - Ay), gp.dr[1]));
            for (int i = 0; i < padx; ++i) T[i] = 0.0;
            T += padx;
            for (int i = 0; i < 2 * Ax; ++i) {
                const double b = __dsub_rn(rec.r[0], __dmul_rn((double)(rec.ir[0] - Ax + i), gp.dr[0]));
                T[i] = exp(-(gp.cxx * b * b + 2.0 * gp.gxy * b * by0) * it2);
            }
            T += 2 * Ax;
            for (int i = 0; i < padx + pady; ++i) T[i] = 0.0;
            T += padx + pady;
            for (int j = 0; j < 2 * Ay; ++j) {
                const double b = __dsub_rn(rec.r[1], __dmul_rn((double)(rec.ir[1] - Ay + j), gp.dr[1]));
                T[j] = exp(-(gp.cyy * b * b - 2.0 * gp.gxy * ((double)j * gp.dr[1]) * bx0) * it2);
            }
            T += 2 * Ay;
            for (int i = 0; i < pady; ++i) T[i] = 0.0;
            T += pady;
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

// ---- K2: fill the lists.  The splat accumulates in 64-bit fixed point (integer adds commute), so the order inside
// a list is irrelevant and a counting sort with atomics is deterministic in its RESULT: the exclusive scan of the K1
// counts gives the list starts, and every pair claims its slot with ONE atomic on that same array
// (pos = atomicAdd(&cur[key], 1)): afterwards cur[key] is the END of list `key`, i.e. the start of list key+1, which is
// all the splat needs (lists are contiguous; list 0 starts at 0).  One record is one aligned 32-byte sector (PairRec +
// PairAux + padding): the lists are written at random positions, and a full-sector store needs no fill read from DRAM
// (16 + 8-byte stores into two arrays plus a separate start / cursor pair moved 130 B of random DRAM sectors per pair:
// the kernel was bound by those, not by the atomics).  Records and atom records stream (evict-first) so that the
// per-list array stays L2-resident.
// ---- K2a: walk order of K2.  The lists of one frame are ~200 MB of 32-byte records at c3; atoms in file order hit
// them at random, every record store then costs DRAM a 64-byte read-modify-write (ncu: 117 B of traffic per 32-byte
// record).  Atoms ordered by the x tile of their home cell write into a window of a few tile rows -- some MB that
// stay in L2 until whole lines are complete.  Counting sort: xcnt holds the bucket starts (exclusive scan of the K1
// counts) and is advanced by one atomic per atom; the order inside a bucket is irrelevant (see below).
__global__ void __launch_bounds__(256)
order_atoms_kernel(const AtomRec* __restrict__ recs, unsigned* __restrict__ xcnt, unsigned* __restrict__ perm,
                   GridParams gp, int nframes)
{
    const long long total = (long long)nframes * gp.natoms;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(idx / gp.natoms);
        const int ir0 = recs[idx].ir[0];
        perm[atomicAdd(xcnt + f * gp.ntx + x_bucket(ir0, gp), 1u)] = (unsigned)idx;
    }
}

// K2b: list lengths (RED on count[key]); their exclusive scan gives the list starts
__global__ void __launch_bounds__(256)
bin_count_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ perm, unsigned* __restrict__ count,
                 GridParams gp, TypeTable tt, int nframes)
{
    __shared__ __align__(16) AtomRec s_rec[8][32];
    __shared__ unsigned s_idx[8][64];
    const int warp = threadIdx.x >> 5;
    const long long total = (long long)nframes * gp.natoms;
    for (long long slot0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) & ~31LL; slot0 < total;
         slot0 += (long long)gridDim.x * blockDim.x)
        walk_pairs_warp(recs, perm, slot0, total, gp, tt, s_rec[warp], s_idx[warp],
                        [&](unsigned key, const PairRec&, const PairAux&) { atomicAdd(count + key, 1u); });
}

__global__ void __launch_bounds__(256)
bin_place_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ perm /* walk order (nullptr: atom order) */,
                 unsigned* __restrict__ cur /* in: list starts; out: list ends */,
                 uint4* __restrict__ prec2 /* [2 * pairs] */, GridParams gp, TypeTable tt, int nframes,
                 unsigned nkeys, unsigned long long cap, int* __restrict__ err_flag)
{
    __shared__ __align__(16) AtomRec s_rec[8][32];
    __shared__ unsigned s_idx[8][64];
    if ((unsigned long long)cur[nkeys] > cap) {         // total pairs (no list has index nkeys, so nobody moves this entry); cannot
                                                        // exceed the capacity unless the host bound is wrong: refuse to overrun
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(err_flag, 4);
        return;
    }
    const int warp = threadIdx.x >> 5;
    const long long total = (long long)nframes * gp.natoms;
    for (long long slot0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) & ~31LL; slot0 < total;
         slot0 += (long long)gridDim.x * blockDim.x)
        walk_pairs_warp(recs, perm, slot0, total, gp, tt, s_rec[warp], s_idx[warp],
                        [&](unsigned key, const PairRec& pr, const PairAux& pa) {
                            const unsigned pos = atomicAdd(cur + key, 1u);
                            __stcs(prec2 + 2 * (size_t)pos, pr);
                            __stcs(prec2 + 2 * (size_t)pos + 1, make_uint4(pa.x, pa.y, 0u, 0u));
                        });
}
