// K3: deterministic owner-computes Gaussian splat of a column tile, fused with the z-axis FFT.
//
// Replaces the per-atom Python loop dens.py:283-308 and the 26-region fold dens.py:86-108.
// A CTA owns tx*ty (x,y) columns over all z for ONE pair of frames (frame 2q -> real part,
// frame 2q+1 -> imaginary part; threads 0..127 serve the real part, 128..255 the imaginary
// part).  Every thread OWNS the cells of one column inside one z slab; nobody else ever writes
// them, so the accumulation needs no atomics and no barriers, and its order is the order of the
// sorted pair list: the density is bitwise reproducible.  The tile's list (K2, sorted by slab) is
// consumed in chunks:
//   A  one thread per pair: clip the atom image against the tile, fetch the slice of the atom's
//      one-dimensional Gaussian factor tables (built once per atom by K1) that the tile needs into
//      shared memory, publish 16 bytes of geometry; a ballot transpose turns the per-pair column
//      masks into per-column hit lists that keep list order;
//   B  every owner walks the hits of ITS slab sub-list that cover ITS column and adds
//      EX[i]*EY[j]*C[i][j]*EZ[k] into its cells.  The periodic fold (incl. the corner rule of
//      dens.py:107) is an index shift per z segment; no padded array exists.
// Afterwards the tile is transformed along z in place (native FFT path) and stored.
#pragma once
#include "mdsf_common.cuh"
#include "mdsf_fft.cuh"

#ifndef MDSF_SPLAT_MINBLOCKS
#define MDSF_SPLAT_MINBLOCKS 2
#endif
#define MDSF_OWNERS 128            // owner threads per part = columns * slabs
#define MDSF_MAX_STAMP 1023        // 2*A must fit 10 bits of the packed pair info

// what phase B needs to know about a pair, one 16-byte shared-memory load
//   x: cx0 | w << 8 | cy0 << 16 | h << 24          tile-relative clip rectangle
//   y: pz0                                          padded-grid index of the first z cell
//   z: kA | kB << 10 | nz << 20 | lowB << 30 | highB << 31   z segments; *B: shift is Nborder, not N_z
//   w: i0 | j0 << 10 | type << 20                   stamp index of the first clipped column (cross-term table)
typedef int4 PairInfo;

// 64-bit fixed-point add into shared memory from two NATIVE 32-bit atomics (a 64-bit shared-memory
// atomicAdd compiles to a compare-and-swap loop).  The low-word add returns the old value, so exactly the
// adds that wrap the low word see a carry; the high word receives hi + carry.  The final 64-bit value is
// the exact integer sum whatever the interleaving, i.e. still bitwise deterministic.  v must be >= 0.
__device__ __forceinline__ void smem_add_u64(unsigned long long* cell, unsigned long long v) {
    unsigned* w = reinterpret_cast<unsigned*>(cell);
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(w, lo);
    const unsigned carry = (old + lo) < lo ? 1u : 0u;
    if (hi | carry) atomicAdd(w + 1, hi + carry);
}

__device__ __forceinline__ void part_barrier(int part) {
    asm volatile("bar.sync %0, %1;" ::"r"(part + 1), "r"(128) : "memory");
}

// ATOMIC = true is the "tile" splat mode: phase B has no owners, one thread per (pair, column) adds its
// terms with 64-bit FIXED-POINT integer atomics into the shared-memory tile (integer addition commutes, so
// the result is still bitwise deterministic); the tile is converted to fp64 in place before the z FFT.
// EZG = true compiles the path for stamps taller than `zstage` (EZ read from global memory in phase B).
// PREC = true (tile mode with direct binning): the lists hold 16-byte PAIR RECORDS written by bin_pairs_kernel
// (cell index, type, table offset, image) instead of 4-byte payloads that point at 48-byte atom records: one
// dependent global round trip less per tile, a third of the bytes, contiguous instead of gathered.
//   x: (ir_x + 1024) | (ir_y + 1024) << 12 | (sx + 1) << 24 | (sy + 1) << 26
//   y: (ir_z + 1024) | type << 13          z: table offset          w: atom index
// T44 = true compiles the common geometry in: 4x4-column tiles with 16-entry table slots (tx, ty, the slot stride and
// every shift derived from them become constants; phase A's staging loop drops from ~28 to ~4 instructions per slot).
template <bool FUSE_ZFFT, bool ATOMIC, bool EZG, bool PREC, bool T44>
__global__ void __launch_bounds__(256, MDSF_SPLAT_MINBLOCKS)
splat_zfft_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ vals,
                  const unsigned* __restrict__ tile_start, double2* __restrict__ vol,
                  double2* __restrict__ dens_dump, GridParams gp, TypeTable tt, FftPlan zplan,
                  const double2* __restrict__ twz, const double* __restrict__ atom_tables, int chunk, int logS, int zfast,
                  int zstage, int* __restrict__ err_flag, const double2* __restrict__ tw16, int tw16_off, int pf_dist)
{
    extern __shared__ double smem[];
    // Nz = 256 tile mode: stage-1 twiddles of the 16x16 z split arrive by cp.async while the splat runs
    double2* tw16s = reinterpret_cast<double2*>(reinterpret_cast<char*>(smem) + tw16_off);
    if (FUSE_ZFFT && ATOMIC && tw16_off > 0) cp_async16(tw16s + threadIdx.x, tw16 + threadIdx.x);
    const int TX = T44 ? 4 : gp.tx, TY = T44 ? 4 : gp.ty;
    if (T44) logS = 4;
    const int ncol = TX * TY;
    const int nzp = gp.nzp;
    const int ntiles = gp.ntx * gp.nty;
    const int tile = blockIdx.x, q = blockIdx.y;
    const int X0 = (tile / gp.nty) * TX, Y0 = (tile % gp.nty) * TY;
    const int part = threadIdx.x >> 7, pt = threadIdx.x & 127, lane = threadIdx.x & 31, pw = pt >> 5;
    const int nz = gp.n[2];
    const int f = 2 * q + part;

    // ---- ownership: owner o = slab * ncol + column; slab s covers z in [s*zs, (s+1)*zs).
    // The tile's pair list is sorted by slab (K2), so an owner's hits are the pairs of ITS slab
    // sub-list that cover its column -- no filtering, every loop iteration does real work.
    const int nslab = gp.nslab, zs = gp.zs;
    const int mycol = pt % ncol, myslab = ATOMIC ? 0 : pt / ncol;
    const unsigned kbase = (unsigned)(f * ntiles + tile) * (unsigned)nslab;
    const unsigned lbeg = tile_start[kbase], lend = (gp.debug_skip & 16) ? lbeg : tile_start[kbase + nslab];
    // L2 warm-up for the CTA that runs `pf_dist` tiles later (PREC lists): its list bounds are requested together with
    // ours, its first pair records together with ours (the load itself pulls them into L2), and its atoms' factor tables
    // are prefetched into L2 once those records have arrived -- the later CTA's dependent loads then hit L2, not HBM
    unsigned fbeg = 0, fend = 0;
    if (PREC && pf_dist > 0) {
        const long long lin = (long long)q * ntiles + tile + pf_dist;
        const int qf = (int)(lin / ntiles), tf = (int)(lin - (long long)qf * ntiles);
        if (qf < (int)gridDim.y) {
            const unsigned kf = (unsigned)((2 * qf + part) * ntiles + tf);
            fbeg = tile_start[kf]; fend = tile_start[kf + 1];
        }
    }
    const unsigned sbeg = tile_start[kbase + myslab], send = tile_start[kbase + myslab + 1];

    // ---- shared memory carve-up
    double* tile_re = smem;                                   // [ncol][nzp]  frame 2q
    double* tile_im = tile_re + (size_t)ncol * nzp;           // [ncol][nzp]  frame 2q+1
    double* tables = tile_im + (size_t)ncol * nzp;
    const size_t tbl_per_part = (size_t)chunk << logS;        // per pair: [EX: tx][EY: ty][EZ: 2Az] in a 2^logS stride
    double* tbl = tables + part * tbl_per_part;
    double* rxyz = tbl;                                       // general-ucell path keeps r here instead (3 doubles per pair)
    double* twr = tables;                                     // z twiddles reuse the table region once the splat is done
    double* twi = twr + nz;
    PairInfo* info_all = reinterpret_cast<PairInfo*>(tables + 2 * tbl_per_part);
    PairInfo* info = info_all + part * chunk;
    unsigned* hitT = reinterpret_cast<unsigned*>(info_all + 2 * chunk) + part * 4 * 32;   // [4 warps of pairs][column]
    int* voff = reinterpret_cast<int*>(reinterpret_cast<unsigned*>(info_all + 2 * chunk) + 2 * 4 * 32) + part * (chunk + 8);   // ATOMIC: visit offsets
    unsigned char* vpair = reinterpret_cast<unsigned char*>(reinterpret_cast<int*>(reinterpret_cast<unsigned*>(info_all + 2 * chunk) + 2 * 4 * 32) + 2 * (chunk + 8))
                           + (size_t)part * chunk * ncol;      // ATOMIC: visit -> pair (chunk <= 128 pairs, <= ncol visits each)

    double* mytile = part ? tile_im : tile_re;
    // the first chunk's list entries and atom records are requested before the tile is cleared, the next
    // chunk's while the current one is being accumulated: the dependent loads overlap useful work
    unsigned pf_v = 0;
    AtomRec pf_rec;
    pf_rec.type = 0;
    uint4 pf_p = make_uint4(0u, 0u, 0u, 0u);
    const uint4* __restrict__ prec = reinterpret_cast<const uint4*>(vals);
    unsigned fut_tbase = 0xffffffffu;
    if (PREC && pt < (int)min((unsigned)chunk, fend - fbeg)) fut_tbase = prec[fbeg + pt].z;
    if (lbeg < lend && pt < (int)min((unsigned)chunk, lend - lbeg)) {
        if (PREC) {
            pf_p = prec[lbeg + pt];
        } else {
            pf_v = vals[lbeg + pt];
            pf_rec = recs[(long long)f * gp.natoms + (int)(pf_v & (MDSF_MAX_ATOMS - 1))];
        }
    }
    {
        double2* z2 = reinterpret_cast<double2*>(tile_re);
        for (int i = threadIdx.x; i < ncol * nzp; i += blockDim.x) z2[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (PREC && fut_tbase != 0xffffffffu) {
        const double* T = atom_tables + fut_tbase;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(T));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(T + 15));
    }

    const int lty = T44 ? 2 : __ffs(TY) - 1;                         // tx, ty are powers of two
    const int mycx = mycol >> lty, mycy = mycol & (TY - 1);
    const int zlo = myslab * zs, zhi = min(zlo + zs, nz);
    const bool owner_valid = X0 + mycx < gp.n[0] && Y0 + mycy < gp.n[1] && zlo < zhi;
    double* col = mytile + (size_t)mycol * nzp;
    const int S = 1 << logS;

    for (unsigned cb = lbeg; cb < lend; cb += chunk) {
        const int npair = (int)min((unsigned)chunk, lend - cb);
        // ---------------- A: one pair per thread
        unsigned colmask = 0;
        if (pt < npair) {
            int sx, sy;
            struct { int ir[3]; int type; unsigned tbase; } rec;
            unsigned atom;
            if (PREC) {
                const uint4 p = pf_p;
                rec.ir[0] = (int)(p.x & 4095u) - 1024; rec.ir[1] = (int)((p.x >> 12) & 4095u) - 1024;
                sx = (int)((p.x >> 24) & 3u) - 1; sy = (int)((p.x >> 26) & 3u) - 1;
                rec.ir[2] = (int)(p.y & 8191u) - 1024; rec.type = (int)(p.y >> 13);
                rec.tbase = p.z; atom = p.w;
            } else {
                const unsigned v = pf_v;
                sx = (int)((v >> MDSF_ATOM_BITS) & 3u) - 1; sy = (int)((v >> (MDSF_ATOM_BITS + 2)) & 3u) - 1;
                rec.ir[0] = pf_rec.ir[0]; rec.ir[1] = pf_rec.ir[1]; rec.ir[2] = pf_rec.ir[2];
                rec.type = pf_rec.type; rec.tbase = pf_rec.tbase; atom = v & (MDSF_MAX_ATOMS - 1);
            }
            const int Ax = tt.halfw[rec.type * 3], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
            int xlo, xhi, ylo, yhi;
            stamp_segment(rec.ir[0], Ax, gp.n[0], sx, xlo, xhi);
            stamp_segment(rec.ir[1], Ay, gp.n[1], sy, ylo, yhi);
            // destination range of this image, clipped to the tile (tile-relative)
            const int cx0 = max(xlo - sx * gp.n[0] - X0, 0), cx1 = min(xhi - sx * gp.n[0] - X0, TX);
            const int cy0 = max(ylo - sy * gp.n[1] - Y0, 0), cy1 = min(yhi - sy * gp.n[1] - Y0, TY);
            const int w = max(cx1 - cx0, 0), h = max(cy1 - cy0, 0);
            const int i0 = X0 + cx0 + sx * gp.n[0] - (rec.ir[0] - Ax);       // stamp index of the first clipped column
            const int j0 = Y0 + cy0 + sy * gp.n[1] - (rec.ir[1] - Ay);
            int shlo, shhi, kA, kB;
            (void)image_slabmask(rec.ir[2], Az, sx, sy, nz, gp.nb, gp.fold_mode, zs, shlo, shhi, kA, kB);
            PairInfo pi;
            pi.x = cx0 | (w << 8) | (cy0 << 16) | (h << 24);
            pi.y = rec.ir[2] - Az;
            pi.z = kA | (kB << 10) | ((2 * Az) << 20) | ((shlo != nz) ? (1 << 30) : 0) | ((shhi != -nz) ? (int)(1u << 31) : 0);
            // stamps taller than zstage cells do not stage EZ: phase B reads it from the atom's table in global memory
            const bool ez_global = EZG && 2 * Az > zstage;
            pi.w = i0 | (j0 << 10) | (rec.type << 20) | (ez_global ? (int)(1u << 31) : 0);
            info[pt] = pi;
            if (w * h > 0)
                for (int cx = cx0; cx < cx0 + w; ++cx)
                    colmask |= (((h >= 32) ? 0xffffffffu : ((1u << h) - 1u)) << (cx * TY + cy0));
            if (gp.separable) {
                // slice of the atom's factor tables this tile needs: EX[i0..i0+w), EY[j0..j0+h), EZ[0..2Az)
                if (!(gp.debug_skip & 4)) {
                    const double* T = atom_tables + rec.tbase;
                    double* dst = tbl + pt;                 // transposed: entry sidx of pair pt at tbl[sidx*chunk + pt] (conflict-free)
                    if (S <= 16) {              // issue every load before the first store
                        double vv[16];
#pragma unroll
                        for (int sidx = 0; sidx < 16; ++sidx) {
                            int src = -1;
                            if (sidx < TX) { if (sidx < w) src = i0 + sidx; }
                            else if (sidx < TX + TY) { if (sidx - TX < h) src = 2 * Ax + j0 + (sidx - TX); }
                            else if (!ez_global && sidx - TX - TY < 2 * Az) src = 2 * (Ax + Ay) + (sidx - TX - TY);
                            vv[sidx] = (src >= 0 && sidx < S) ? T[src] : 0.0;
                        }
                        if (ez_global) vv[TX + TY] = __longlong_as_double((long long)rec.tbase + 2 * (Ax + Ay));
#pragma unroll
                        for (int sidx = 0; sidx < 16; ++sidx) if (sidx < S) dst[sidx * chunk] = vv[sidx];
                    } else {
                        for (int sidx = 0; sidx < w; ++sidx) dst[sidx * chunk] = T[i0 + sidx];
                        for (int sidx = 0; sidx < h; ++sidx) dst[(TX + sidx) * chunk] = T[2 * Ax + j0 + sidx];
                        if (ez_global) dst[(TX + TY) * chunk] = __longlong_as_double((long long)rec.tbase + 2 * (Ax + Ay));
                        else
#pragma unroll 4
                            for (int sidx = 0; sidx < 2 * Az; ++sidx) dst[(TX + TY + sidx) * chunk] = T[2 * (Ax + Ay) + sidx];
                    }
                }
            } else {
                const double* rr = PREC ? recs[(long long)f * gp.natoms + (int)atom].r : pf_rec.r;     // general ucell: the coordinate itself
                rxyz[pt * 3] = rr[0]; rxyz[pt * 3 + 1] = rr[1]; rxyz[pt * 3 + 2] = rr[2];
            }
        }
        int nvis = 0;
        if (ATOMIC) {
            // inclusive scan of the pairs' column counts over the part's 128 threads -> voff[i+1]
            nvis = __popc(colmask);
            int v = nvis;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t0 = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t0; }
            if (lane == 31) voff[chunk + 1 + pw] = v;
            part_barrier(part);
            int basev = 0;
            for (int w2 = 0; w2 < pw; ++w2) basev += voff[chunk + 1 + w2];
            if (pt < chunk) voff[pt + 1] = v + basev;
            if (pt == 0) voff[0] = 0;
            for (int u = 0; u < nvis; ++u) vpair[v + basev - nvis + u] = (unsigned char)pt;
        } else {
            // ballot transpose of the 32 x ncol (pair x column) hit matrix of this warp:
            // hitT[warp][c] = pairs (bit = lane) whose image covers column c; bit order = list order
            unsigned mine = 0;
            for (int c = 0; c < ncol; ++c) {
                const unsigned b = __ballot_sync(0xffffffffu, (colmask >> c) & 1u);
                if (lane == c) mine = b;
            }
            hitT[pw * 32 + lane] = mine;
        }
        part_barrier(part);
        if (cb + chunk < lend && pt < (int)min((unsigned)chunk, lend - cb - chunk)) {      // prefetch the next chunk
            if (PREC) {
                pf_p = prec[cb + chunk + pt];
            } else {
                pf_v = vals[cb + chunk + pt];
                pf_rec = recs[(long long)f * gp.natoms + (int)(pf_v & (MDSF_MAX_ATOMS - 1))];
            }
        }

        // ---------------- B (tile mode): one thread per (pair, column), fixed-point integer atomics
        if (ATOMIC && !(gp.debug_skip & 1)) {
            unsigned long long* itile = reinterpret_cast<unsigned long long*>(mytile);
            const int total = voff[npair];
            for (int v = pt; v < total; v += 128) {
                const int lo = vpair[v];                  // pair i with voff[i] <= v < voff[i+1]
                const PairInfo pi = info[lo];
                const int hh = (pi.x >> 24) & 0xff;
                const int local = v - voff[lo];
                // T44: hh <= 4 and local < 16, so the division is an exact multiply-shift
                const int lx = T44 ? (int)(((unsigned)local * (hh == 1 ? 65536u : (hh == 2 ? 32768u : (hh == 3 ? 21846u : 16384u)))) >> 16) : local / hh;
                const int ly = local - lx * hh;
                const int cx = (pi.x & 0xff) + lx, cy = ((pi.x >> 16) & 0xff) + ly;
                if (X0 + cx >= gp.n[0] || Y0 + cy >= gp.n[1]) continue;
                const int pz0 = pi.y, kA = pi.z & 1023, kB = (pi.z >> 10) & 1023, nzr = (pi.z >> 20) & 1023;
                const int shlo = (pi.z & (1 << 30)) ? gp.nb : nz, shhi = (pi.z < 0) ? -gp.nb : -nz;
                unsigned long long* colp = itile + (size_t)((cx << lty) + cy) * nzp;
                if (gp.separable) {
                    const double* T = tbl + lo;
                    double exy = T[lx * chunk] * T[(TX + ly) * chunk];
                    if (tt.ctab != nullptr) {
                        const int type = (pi.w >> 20) & 1023, i0 = pi.w & 1023, j0 = (pi.w >> 10) & 1023;
                        exy *= tt.ctab[tt.ctab_off[type] + (i0 + lx) * 2 * tt.halfw[type * 3 + 1] + (j0 + ly)];
                    }
                    exy *= gp.fx_scale;
                    const double* ez = T + (TX + TY) * chunk;
                    int ezs = chunk;                          // stride of the EZ entries (compile-time chunk stride when !EZG)
                    if (EZG && pi.w < 0) { ez = atom_tables + __double_as_longlong(ez[0]); ezs = 1; }
                    for (int k = 0; k < nzr; ++k) {
                        const int pz = pz0 + k;
                        const int cz = k < kA ? pz + shlo : (k < kB ? pz : pz + shhi);
                        smem_add_u64(colp + cz + (cz >> gp.pad_shift), (unsigned long long)__double2ll_rn(exy * ez[k * ezs]));
                    }
                } else {
                    const int type = (pi.w >> 20) & 1023, i0 = pi.w & 1023, j0 = (pi.w >> 10) & 1023;
                    const int Ax = tt.halfw[type * 3], Ay = tt.halfw[type * 3 + 1];
                    const double rx = rxyz[lo * 3], ry = rxyz[lo * 3 + 1], rz = rxyz[lo * 3 + 2];
                    const int px = (int)(rx / gp.dr[0]) - Ax + i0 + lx, py = (int)(ry / gp.dr[1]) - Ay + j0 + ly;
                    const double bx = __dsub_rn(rx, __dmul_rn((double)px, gp.dr[0]));
                    const double by = __dsub_rn(ry, __dmul_rn((double)py, gp.dr[1]));
                    const double t2 = tt.two_sig2[type], amp = tt.amp[type] * gp.fx_scale;
                    for (int k = 0; k < nzr; ++k) {
                        const int pz = pz0 + k;
                        const int cz = k < kA ? pz + shlo : (k < kB ? pz : pz + shhi);
                        const double bzv = __dsub_rn(rz, __dmul_rn((double)pz, gp.dr[2]));
                        const double c0 = gp.u[0] * bx + gp.u[3] * by + gp.u[6] * bzv;
                        const double c1 = gp.u[1] * bx + gp.u[4] * by + gp.u[7] * bzv;
                        const double c2 = gp.u[2] * bx + gp.u[5] * by + gp.u[8] * bzv;
                        smem_add_u64(colp + cz + (cz >> gp.pad_shift), (unsigned long long)__double2ll_rn(amp * exp(-(c0 * c0 + c1 * c1 + c2 * c2) / t2)));
                    }
                }
            }
        }
        // ---------------- B: every owner adds its hits, in list order, into cells only it writes
        if (!ATOMIC && owner_valid && !(gp.debug_skip & 1)) {
            // my slab's pairs inside this chunk are [r0, r1); every owner starts at ITS first warp-of-pairs, so
            // the lanes of a warp (two slabs x 16 columns) all have work in the same loop iteration
            const int r0 = max((int)(sbeg - cb), 0), r1 = min((int)(send - cb), npair);
            for (int wv = r0 >> 5; wv <= (r1 - 1) >> 5 && r1 > r0; ++wv) {
                const int lo = max(r0 - wv * 32, 0), hi = min(r1 - wv * 32, 32);
                unsigned m = hitT[wv * 32 + mycol] & (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                while (m) {
                    const int i = wv * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const PairInfo pi = info[i];
                    const int lx = mycx - (pi.x & 0xff), ly = mycy - ((pi.x >> 16) & 0xff);
                    const int pz0 = pi.y, kA = pi.z & 1023, kB = (pi.z >> 10) & 1023, nzr = (pi.z >> 20) & 1023;
                    const int shlo = (pi.z & (1 << 30)) ? gp.nb : nz, shhi = (pi.z < 0) ? -gp.nb : -nz;
                    if (gp.separable) {
                        const double* T = tbl + i;
                        double exy = T[lx * chunk] * T[(TX + ly) * chunk];
                        if (tt.ctab != nullptr) {
                            const int type = (pi.w >> 20) & 1023, i0 = pi.w & 1023, j0 = (pi.w >> 10) & 1023;
                            exy *= tt.ctab[tt.ctab_off[type] + (i0 + lx) * 2 * tt.halfw[type * 3 + 1] + (j0 + ly)];
                        }
                        const double* ez = T + (TX + TY) * chunk;
                        int ezs = chunk;
                        if (EZG && pi.w < 0) { ez = atom_tables + __double_as_longlong(ez[0]); ezs = 1; }
                        {   // cell
                            const int ka = max(kA, zlo - pz0), kb = min(kB, zhi - pz0);
                            for (int k = ka; k < kb; ++k) { const int cz = pz0 + k; col[cz + (cz >> gp.pad_shift)] += exy * ez[k * ezs]; }
                        }
                        if (kA > 0) {   // low padding
                            const int sh = pz0 + shlo, ka = max(0, zlo - sh), kb = min(kA, zhi - sh);
                            for (int k = ka; k < kb; ++k) { const int cz = sh + k; col[cz + (cz >> gp.pad_shift)] += exy * ez[k * ezs]; }
                        }
                        if (nzr > kB) { // high padding
                            const int sh = pz0 + shhi, ka = max(kB, zlo - sh), kb = min(nzr, zhi - sh);
                            for (int k = ka; k < kb; ++k) { const int cz = sh + k; col[cz + (cz >> gp.pad_shift)] += exy * ez[k * ezs]; }
                        }
                    } else {
                        // general ucell: one exp per cell, exactly the reference's expression
                        const int type = (pi.w >> 20) & 1023, i0 = pi.w & 1023, j0 = (pi.w >> 10) & 1023;
                        const int Ax = tt.halfw[type * 3], Ay = tt.halfw[type * 3 + 1];
                        const double rx = rxyz[i * 3], ry = rxyz[i * 3 + 1], rz = rxyz[i * 3 + 2];
                        // padded-grid index of my column: p = ir - A + stamp index, ir = trunc(r/dr) as in K1
                        const int px = (int)(rx / gp.dr[0]) - Ax + i0 + lx, py = (int)(ry / gp.dr[1]) - Ay + j0 + ly;
                        const double bx = __dsub_rn(rx, __dmul_rn((double)px, gp.dr[0]));
                        const double by = __dsub_rn(ry, __dmul_rn((double)py, gp.dr[1]));
                        const double t2 = tt.two_sig2[type], amp = tt.amp[type];
#pragma unroll 1
                        for (int seg = 0; seg < 3; ++seg) {
                            const int sh = pz0 + (seg == 0 ? 0 : (seg == 1 ? shlo : shhi));
                            const int k0 = seg == 0 ? kA : (seg == 1 ? 0 : kB), k1 = seg == 0 ? kB : (seg == 1 ? kA : nzr);
                            const int ka = max(k0, zlo - sh), kb = min(k1, zhi - sh);
                            for (int k = ka; k < kb; ++k) {
                                const double bzv = __dsub_rn(rz, __dmul_rn((double)(pz0 + k), gp.dr[2]));
                                const double c0 = gp.u[0] * bx + gp.u[3] * by + gp.u[6] * bzv;
                                const double c1 = gp.u[1] * bx + gp.u[4] * by + gp.u[7] * bzv;
                                const double c2 = gp.u[2] * bx + gp.u[5] * by + gp.u[8] * bzv;
                                const int cz = sh + k;
                                col[cz + (cz >> gp.pad_shift)] += amp * exp(-(c0 * c0 + c1 * c1 + c2 * c2) / t2);
                            }
                        }
                    }
                }
            }
        }
        part_barrier(part);
    }
    __syncthreads();
    // tile mode on the fast z path converts fixed point -> fp64 inside the first FFT stage instead
    const bool fused_convert = ATOMIC && FUSE_ZFFT && zfast && dens_dump == nullptr && !(gp.debug_skip & 2);
    if (ATOMIC && !fused_convert) {   // fixed point -> fp64, in place (overflow: |value| >= 2^62: a cell held > 2048 peak amplitudes)
        long long ovf = 0;
        for (int i = threadIdx.x; i < 2 * ncol * nzp; i += blockDim.x) {
            const long long q64 = reinterpret_cast<long long*>(tile_re)[i];
            ovf |= q64 ^ (q64 << 1);
            tile_re[i] = (double)q64 * gp.fx_inv;
        }
        if (ovf < 0) atomicExch(err_flag, 2);
    }
    const bool tw_ready = FUSE_ZFFT && ATOMIC && tw16_off > 0 && fused_convert && zfast == 16;
    if (FUSE_ZFFT && ATOMIC && tw16_off > 0) cp_async_wait_all();
    if (FUSE_ZFFT && !tw_ready) {
        if (fused_convert && zfast == 16) {
            // stage-1 twiddles laid out [k][n2] (w^(n2 k) at k*16 + n2): lanes walk n2, so the reads are conflict-free
            for (int i = threadIdx.x; i < 256; i += blockDim.x) {
                const double2 w = twz[(i >> 4) * (i & 15)];
                twr[i] = w.x; twi[i] = w.y;
            }
        } else {
            load_twiddles(twr, twi, twz, nz);
        }
    }
    if (ATOMIC || FUSE_ZFFT) __syncthreads();

    if (dens_dump != nullptr) {
        for (int i = threadIdx.x; i < ncol * nz; i += blockDim.x) {
            const int c = i / nz, z = i - c * nz;
            const int x = X0 + (c >> lty), y = Y0 + (c & (TY - 1));
            if (x < gp.n[0] && y < gp.n[1]) {
                const int a = c * nzp + z + (z >> gp.pad_shift);
                dens_dump[(((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz + z] = make_double2(tile_re[a], tile_im[a]);
            }
        }
        __syncthreads();     // the FFT below rewrites the tile in place
    }
    if (FUSE_ZFFT && zfast && !(gp.debug_skip & 2)) {
        // Nz = R*R: first radix-R stage in place in shared memory, second stage from shared memory straight
        // to the volume.  Output k2 of block k1 goes to z = R*k2 + k1: the transposed digit-reversed order of
        // two equal radices is the natural frequency order, and lanes (k1) write contiguous 16-byte cells.
        const int R = zfast;
        if (fused_convert && R == 16) {
            // stage 1 written out: load the int64 cells, check overflow, convert, radix-16, twiddle, store fp64 in place
            long long ovf = 0;
            for (int it = threadIdx.x; it < ncol * 16; it += blockDim.x) {
                const int fcol = it >> 4, n2 = it & 15;
                double xr[16], xi[16];
                int addr[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int p = n2 + 16 * j;
                    addr[j] = fcol * nzp + p + (p >> gp.pad_shift);
                    const long long qr = reinterpret_cast<long long*>(tile_re)[addr[j]], qi = reinterpret_cast<long long*>(tile_im)[addr[j]];
                    ovf |= (qr ^ (qr << 1)) | (qi ^ (qi << 1));
                    xr[j] = (double)qr * gp.fx_inv; xi[j] = (double)qi * gp.fx_inv;
                }
                Dft<16>::run(xr, xi, twr, twi, nz);
#pragma unroll
                for (int k = 1; k < 16; ++k) {
                    double wr, wi;
                    if (tw_ready) { const double2 w = tw16s[k * 16 + n2]; wr = w.x; wi = w.y; }
                    else { wr = twr[k * 16 + n2]; wi = twi[k * 16 + n2]; }
                    const double yr = xr[k] * wr - xi[k] * wi;
                    xi[k] = xr[k] * wi + xi[k] * wr;
                    xr[k] = yr;
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) { tile_re[addr[k]] = xr[k]; tile_im[addr[k]] = xi[k]; }
            }
            if (ovf < 0) atomicExch(err_flag, 2);
        } else {
            if (fused_convert) {      // radix 8: plain conversion pass, then the generic stage
                long long ovf = 0;
                for (int i = threadIdx.x; i < 2 * ncol * nzp; i += blockDim.x) {
                    const long long q64 = reinterpret_cast<long long*>(tile_re)[i];
                    ovf |= q64 ^ (q64 << 1);
                    tile_re[i] = (double)q64 * gp.fx_inv;
                }
                if (ovf < 0) atomicExch(err_flag, 2);
                __syncthreads();
            }
            fft_stage_dispatch<true, IO_SMEM, IO_SMEM>(R, tile_re, tile_im, twr, twi, nz, nz, ncol, nzp, 1, gp.pad_shift, 0,
                                                       GlobalTile{nullptr, 0, 0}, nullptr, false);
        }
        __syncthreads();
        for (int it = threadIdx.x; it < ncol * R; it += blockDim.x) {
            const int fcol = it / R, b = it - fcol * R;
            const int x = X0 + (fcol >> lty), y = Y0 + (fcol & (TY - 1));
            if (x >= gp.n[0] || y >= gp.n[1]) continue;
            double2* dst = vol + (((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz + b;
            if (R == 16) {
                double xr[16], xi[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { const int p = b * 16 + j; const int a = fcol * nzp + p + (p >> gp.pad_shift); xr[j] = tile_re[a]; xi[j] = tile_im[a]; }
                Dft<16>::run(xr, xi, twr, twi, nz);
#pragma unroll
                for (int k = 0; k < 16; ++k) dst[16 * k] = make_double2(xr[k], xi[k]);
            } else {
                double xr[8], xi[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int p = b * 8 + j; const int a = fcol * nzp + p + (p >> gp.pad_shift); xr[j] = tile_re[a]; xi[j] = tile_im[a]; }
                Dft<8>::run(xr, xi, twr, twi, nz);
#pragma unroll
                for (int k = 0; k < 8; ++k) dst[8 * k] = make_double2(xr[k], xi[k]);
            }
        }
        return;
    }
    if (FUSE_ZFFT && !(gp.debug_skip & 2)) fft_tile_z(tile_re, tile_im, twr, twi, zplan, ncol, nzp, gp.pad_shift);
    if (!(gp.debug_skip & 8)) {
        // thread <-> z, loop over the tile's columns: 16-byte stores, one contiguous run per column
        for (int z = threadIdx.x; z < nz; z += blockDim.x) {
            const int a0 = z + (z >> gp.pad_shift);
            const int ymax = min(TY, gp.n[1] - Y0);
            for (int cx = 0; cx < TX; ++cx) {
                const int x = X0 + cx;
                if (x >= gp.n[0]) break;
                double2* dst = vol + (((long long)q * gp.n[0] + x) * gp.n[1] + Y0) * nz + z;
                int a = (cx << lty) * nzp + a0;
                for (int cy = 0; cy < ymax; ++cy, a += nzp, dst += nz) *dst = make_double2(tile_re[a], tile_im[a]);
            }
        }
    }
}
