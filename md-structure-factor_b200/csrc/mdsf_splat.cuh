// K3: deterministic owner-computes Gaussian splat of a column tile, fused with the z-axis FFT.
//
// Replaces the per-atom Python loop dens.py:283-308 and the 26-region fold dens.py:86-108.
// A CTA owns tx*ty (x,y) columns over all z for ONE pair of frames (frame 2q -> real part,
// frame 2q+1 -> imaginary part; threads 0..127 serve the real part, 128..255 the imaginary
// part).  Every thread OWNS the cells of one column inside one z slab; nobody else ever writes
// them, so the accumulation needs no atomics and no barriers, and its order is the order of the
// sorted pair list: the density is bitwise reproducible.  The list (K2) is consumed in chunks:
//   A0  one thread per pair: clip the atom image against the tile, work out which owners
//       (column x slab) it touches; a ballot transpose turns the per-pair owner masks into
//       per-owner hit lists that keep list order;
//   A1  the Gaussian factors are evaluated densely, one table entry per thread:
//       exy[pair][column] = exp(-(c0^2+c1^2)/(2 sigma^2)),  ez[pair][k] = Nel/sigma^3 exp(-c2^2/(2 sigma^2));
//   B   every owner walks ITS hits and adds exy*ez[k] into its cells.  The periodic fold (incl.
//       the corner rule of dens.py:107) is an index shift per z segment; no padded array exists.
// Afterwards the tile is transformed along z in place (native FFT path) and stored.
#pragma once
#include "mdsf_common.cuh"
#include "mdsf_fft.cuh"

#ifndef MDSF_SPLAT_MINBLOCKS
#define MDSF_SPLAT_MINBLOCKS 2
#endif
#define MDSF_OWNERS 128            // owner threads per part = columns * slabs

struct PairSlot {
    double rx, ry, rz;     // atom coordinate (general-ucell path only)
    int px0, py0, pz0;     // padded-grid index of the first clipped column / first z cell
    int type;
    unsigned rect;         // cx0 | w << 8 | cy0 << 16 | h << 24   (tile-relative clip rectangle)
    short nz, kA, kB, pad0;// 2*Az; k < kA: low padding, kA <= k < kB: cell, k >= kB: high padding
    int shlo, shhi;        // destination z = pz0 + k + shlo (low padding) / + shhi (high padding)
    unsigned tbase;        // offset of the atom's factor tables (frame block included, in doubles / 1)
    short i0, j0;          // stamp index of the first clipped column (for the cross-term table)
    short twoAx, twoAy;
};

__device__ __forceinline__ void part_barrier(int part) {
    asm volatile("bar.sync %0, %1;" ::"r"(part + 1), "r"(128) : "memory");
}

template <bool FUSE_ZFFT>
__global__ void __launch_bounds__(256, MDSF_SPLAT_MINBLOCKS)
splat_zfft_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ vals,
                  const unsigned* __restrict__ tile_start, double2* __restrict__ vol,
                  double2* __restrict__ dens_dump, GridParams gp, TypeTable tt, FftPlan zplan,
                  const double2* __restrict__ twz, const double* __restrict__ atom_tables, int chunk, int logS)
{
    extern __shared__ double smem[];
    const int ncol = gp.tx * gp.ty;
    const int nzp = gp.nzp;
    const int ntiles = gp.ntx * gp.nty;
    const int tile = blockIdx.x, q = blockIdx.y;
    const int X0 = (tile / gp.nty) * gp.tx, Y0 = (tile % gp.nty) * gp.ty;
    const int part = threadIdx.x >> 7, pt = threadIdx.x & 127, lane = threadIdx.x & 31, pw = pt >> 5;
    const int nz = gp.n[2];

    // ---- shared memory carve-up
    double* tile_re = smem;                                   // [ncol][nzp]  frame 2q
    double* tile_im = tile_re + (size_t)ncol * nzp;           // [ncol][nzp]  frame 2q+1
    double* twr = tile_im + (size_t)ncol * nzp;               // [nz] z twiddles (fused FFT only)
    double* twi = twr + (FUSE_ZFFT ? nz : 0);
    double* tables = twi + (FUSE_ZFFT ? nz : 0);
    const size_t tbl_per_part = (size_t)chunk << logS;        // per pair: [EX: tx][EY: ty][EZ: 2Az] in a 2^logS stride
    double* tbl = tables + part * tbl_per_part;
    PairSlot* slots_all = reinterpret_cast<PairSlot*>(tables + 2 * tbl_per_part);
    unsigned* hit_all = reinterpret_cast<unsigned*>(slots_all + 2 * chunk);
    PairSlot* slots = slots_all + part * chunk;               // [chunk]
    unsigned* hitT = hit_all + part * 4 * 32;                 // [4 warps of pairs][column]

    double* mytile = part ? tile_im : tile_re;
    {
        double2* z2 = reinterpret_cast<double2*>(tile_re);
        for (int i = threadIdx.x; i < ncol * nzp; i += blockDim.x) z2[i] = make_double2(0.0, 0.0);
    }
    if (FUSE_ZFFT) load_twiddles(twr, twi, twz, nz);
    __syncthreads();

    const int f = 2 * q + part;
    // ---- ownership: owner o = slab * ncol + column; slab s covers z in [s*zs, (s+1)*zs).
    // The tile's pair list is sorted by slab (K2), so an owner's hits are the pairs of ITS slab
    // sub-list that cover its column -- no filtering, every loop iteration does real work.
    const int nslab = gp.nslab, zs = gp.zs;
    const int mycol = pt % ncol, myslab = pt / ncol;
    const unsigned kbase = (unsigned)(f * ntiles + tile) * (unsigned)nslab;
    const unsigned lbeg = tile_start[kbase], lend = tile_start[kbase + nslab];
    const unsigned sbeg = tile_start[kbase + myslab], send = tile_start[kbase + myslab + 1];
    const int mycx = mycol / gp.ty, mycy = mycol % gp.ty;
    const int zlo = myslab * zs, zhi = min(zlo + zs, nz);
    const bool owner_valid = X0 + mycx < gp.n[0] && Y0 + mycy < gp.n[1] && zlo < zhi;
    double* col = mytile + (size_t)mycol * nzp;

    for (unsigned cb = lbeg; cb < ((gp.debug_skip & 16) ? lbeg : lend); cb += chunk) {
        const int npair = (int)min((unsigned)chunk, lend - cb);
        // ---------------- A0: clip one pair per thread
        unsigned colmask = 0;
        if (pt < npair) {
            PairSlot s;
            const unsigned v = vals[cb + pt];
            const int a = (int)(v & (MDSF_MAX_ATOMS - 1));
            const int sx = (int)((v >> MDSF_ATOM_BITS) & 3u) - 1, sy = (int)((v >> (MDSF_ATOM_BITS + 2)) & 3u) - 1;
            const AtomRec rec = recs[(long long)f * gp.natoms + a];
            const int Ax = tt.halfw[rec.type * 3], Ay = tt.halfw[rec.type * 3 + 1], Az = tt.halfw[rec.type * 3 + 2];
            int xlo, xhi, ylo, yhi;
            stamp_segment(rec.ir[0], Ax, gp.n[0], sx, xlo, xhi);
            stamp_segment(rec.ir[1], Ay, gp.n[1], sy, ylo, yhi);
            // destination range of this image, clipped to the tile (tile-relative)
            const int cx0 = max(xlo - sx * gp.n[0] - X0, 0), cx1 = min(xhi - sx * gp.n[0] - X0, gp.tx);
            const int cy0 = max(ylo - sy * gp.n[1] - Y0, 0), cy1 = min(yhi - sy * gp.n[1] - Y0, gp.ty);
            const int w = max(cx1 - cx0, 0), h = max(cy1 - cy0, 0);
            s.rx = rec.r[0]; s.ry = rec.r[1]; s.rz = rec.r[2];
            s.px0 = X0 + cx0 + sx * gp.n[0];
            s.py0 = Y0 + cy0 + sy * gp.n[1];
            s.pz0 = rec.ir[2] - Az;
            s.type = rec.type;
            s.rect = (unsigned)cx0 | ((unsigned)w << 8) | ((unsigned)cy0 << 16) | ((unsigned)h << 24);
            int kA, kB;
            (void)image_slabmask(rec.ir[2], Az, sx, sy, nz, gp.nb, gp.fold_mode, zs, s.shlo, s.shhi, kA, kB);
            s.nz = (short)(2 * Az); s.kA = (short)kA; s.kB = (short)kB; s.pad0 = 0;
            s.tbase = (unsigned)((long long)f * gp.tstride + tt.toff[a]);
            s.i0 = (short)(s.px0 - (rec.ir[0] - Ax)); s.j0 = (short)(s.py0 - (rec.ir[1] - Ay));
            s.twoAx = (short)(2 * Ax); s.twoAy = (short)(2 * Ay);
            if (w * h > 0)
                for (int cx = cx0; cx < cx0 + w; ++cx)
                    colmask |= (((h >= 32) ? 0xffffffffu : ((1u << h) - 1u)) << (cx * gp.ty + cy0));
            slots[pt] = s;
        }
        // ballot transpose of the 32 x ncol (pair x column) hit matrix of this warp:
        // hitT[warp][c] = pairs (bit = lane) whose image covers column c; bit order = list order
        {
            unsigned mine = 0;
            for (int c = 0; c < ncol; ++c) {
                const unsigned b = __ballot_sync(0xffffffffu, (colmask >> c) & 1u);
                if (lane == c) mine = b;
            }
            hitT[pw * 32 + lane] = mine;
        }
        part_barrier(part);

        // ---------------- A1: stage the pairs' factor tables (built per atom by K1) in shared memory
        if (gp.separable && !(gp.debug_skip & 4)) {
            const int S = 1 << logS;
            for (int e = pt; e < (npair << logS); e += 128) {
                const int i = e >> logS, sub = e & (S - 1);
                const PairSlot& p = slots[i];
                const int w = (int)((p.rect >> 8) & 0xff), hh = (int)(p.rect >> 24);
                int src = -1;
                if (sub < gp.tx) { if (sub < w) src = p.i0 + sub; }
                else if (sub < gp.tx + gp.ty) { if (sub - gp.tx < hh) src = p.twoAx + p.j0 + (sub - gp.tx); }
                else if (sub - gp.tx - gp.ty < p.nz) src = p.twoAx + p.twoAy + (sub - gp.tx - gp.ty);
                if (src >= 0) tbl[e] = atom_tables[(size_t)p.tbase + src];
            }
            part_barrier(part);
        }

        // ---------------- B: every owner adds its hits, in list order, into cells only it writes
        if (owner_valid && !(gp.debug_skip & 1)) {
            const int nwarp_used = (npair + 31) >> 5;
            for (int wv = 0; wv < nwarp_used; ++wv) {
                // pairs [lo, hi) of this warp-of-pairs belong to my slab
                const int lo = max((int)(sbeg - cb) - wv * 32, 0), hi = min((int)(send - cb) - wv * 32, 32);
                if (hi <= lo) continue;
                unsigned m = hitT[wv * 32 + mycol] & (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                while (m) {
                    const int i = wv * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const PairSlot& p = slots[i];
                    const int cx0 = (int)(p.rect & 0xff), cy0 = (int)((p.rect >> 16) & 0xff);
                    const int lx = mycx - cx0, ly = mycy - cy0;
                    const int pz0 = p.pz0, kA = p.kA, kB = p.kB, nzr = p.nz;
                    if (gp.separable) {
                        const double* T = tbl + ((size_t)i << logS);
                        double exy = T[lx] * T[gp.tx + ly];
                        if (tt.ctab != nullptr) exy *= tt.ctab[tt.ctab_off[p.type] + (p.i0 + lx) * p.twoAy + (p.j0 + ly)];
                        const double* ez = T + gp.tx + gp.ty;
                        {   // cell
                            const int ka = max(kA, zlo - pz0), kb = min(kB, zhi - pz0);
                            for (int k = ka; k < kb; ++k) { const int cz = pz0 + k; col[cz + (cz >> gp.pad_shift)] += exy * ez[k]; }
                        }
                        if (kA > 0) {   // low padding
                            const int sh = pz0 + p.shlo, ka = max(0, zlo - sh), kb = min(kA, zhi - sh);
                            for (int k = ka; k < kb; ++k) { const int cz = sh + k; col[cz + (cz >> gp.pad_shift)] += exy * ez[k]; }
                        }
                        if (nzr > kB) { // high padding
                            const int sh = pz0 + p.shhi, ka = max(kB, zlo - sh), kb = min(nzr, zhi - sh);
                            for (int k = ka; k < kb; ++k) { const int cz = sh + k; col[cz + (cz >> gp.pad_shift)] += exy * ez[k]; }
                        }
                    } else {
                        // general ucell: one exp per cell, exactly the reference's expression
                        const double bx = __dsub_rn(p.rx, __dmul_rn((double)(p.px0 + lx), gp.dr[0]));
                        const double by = __dsub_rn(p.ry, __dmul_rn((double)(p.py0 + ly), gp.dr[1]));
                        const double t2 = tt.two_sig2[p.type], amp = tt.amp[p.type];
#pragma unroll 1
                        for (int seg = 0; seg < 3; ++seg) {
                            const int sh = pz0 + (seg == 0 ? 0 : (seg == 1 ? p.shlo : p.shhi));
                            const int k0 = seg == 0 ? kA : (seg == 1 ? 0 : kB), k1 = seg == 0 ? kB : (seg == 1 ? kA : nzr);
                            const int ka = max(k0, zlo - sh), kb = min(k1, zhi - sh);
                            for (int k = ka; k < kb; ++k) {
                                const double bzv = __dsub_rn(p.rz, __dmul_rn((double)(pz0 + k), gp.dr[2]));
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

    if (dens_dump != nullptr) {
        for (int i = threadIdx.x; i < ncol * nz; i += blockDim.x) {
            const int c = i / nz, z = i - c * nz;
            const int x = X0 + c / gp.ty, y = Y0 + c % gp.ty;
            if (x < gp.n[0] && y < gp.n[1]) {
                const int a = c * nzp + z + (z >> gp.pad_shift);
                dens_dump[(((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz + z] = make_double2(tile_re[a], tile_im[a]);
            }
        }
    }
    if (FUSE_ZFFT && !(gp.debug_skip & 2)) fft_tile_z(tile_re, tile_im, twr, twi, zplan, ncol, nzp, gp.pad_shift);
    const int lty = __ffs(gp.ty) - 1;                     // tx, ty are powers of two
    for (int c = 0; c < ((gp.debug_skip & 8) ? 0 : ncol); ++c) {
        const int x = X0 + (c >> lty), y = Y0 + (c & (gp.ty - 1));
        if (x < gp.n[0] && y < gp.n[1]) {
            double2* dst = vol + (((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz;
            for (int z = threadIdx.x; z < nz; z += blockDim.x) {
                const int a = c * nzp + z + (z >> gp.pad_shift);
                dst[z] = make_double2(tile_re[a], tile_im[a]);
            }
        }
    }
}
