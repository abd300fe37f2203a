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
    double rx, ry, rz;     // atom coordinate (float64 value of the coords dtype)
    int px0, py0, pz0;     // padded-grid index of the first clipped column / first z cell
    int type;
    unsigned rect;         // cx0 | w << 8 | cy0 << 16 | h << 24   (tile-relative clip rectangle)
    short nz, kA, kB, pad0;// 2*Az; k < kA: low padding, kA <= k < kB: cell, k >= kB: high padding
    int shlo, shhi;        // destination z = pz0 + k + shlo (low padding) / + shhi (high padding)
    int offxy, offz;       // table offsets
};

__device__ __forceinline__ void part_barrier(int part) {
    asm volatile("bar.sync %0, %1;" ::"r"(part + 1), "r"(128) : "memory");
}

template <bool FUSE_ZFFT>
__global__ void __launch_bounds__(256, MDSF_SPLAT_MINBLOCKS)
splat_zfft_kernel(const AtomRec* __restrict__ recs, const unsigned* __restrict__ vals,
                  const unsigned* __restrict__ tile_start, double2* __restrict__ vol,
                  double2* __restrict__ dens_dump, GridParams gp, TypeTable tt, FftPlan zplan,
                  const double2* __restrict__ twz, int chunk, int xycap, int zcap)
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
    const size_t tbl_per_part = (size_t)chunk * (xycap + zcap);
    double* tblxy = tables + part * tbl_per_part;             // [chunk*xycap]
    double* tblz = tblxy + (size_t)chunk * xycap;             // [chunk*zcap]
    PairSlot* slots_all = reinterpret_cast<PairSlot*>(tables + 2 * tbl_per_part);
    unsigned* hit_all = reinterpret_cast<unsigned*>(slots_all + 2 * chunk);
    int* scan_all = reinterpret_cast<int*>(hit_all + 2 * 4 * MDSF_OWNERS);
    PairSlot* slots = slots_all + part * chunk;               // [chunk]
    unsigned* hitT = hit_all + part * 4 * MDSF_OWNERS;        // [4 warps of pairs][owner]
    int* scan_tmp = scan_all + part * 16;                     // warp totals + chunk totals

    double* mytile = part ? tile_im : tile_re;
    for (int i = threadIdx.x; i < 2 * ncol * nzp; i += blockDim.x) tile_re[i] = 0.0;
    if (FUSE_ZFFT) load_twiddles(twr, twi, twz, nz);
    __syncthreads();

    const int f = 2 * q + part;
    const unsigned lbeg = tile_start[f * ntiles + tile], lend = tile_start[f * ntiles + tile + 1];

    // ---- ownership: owner o = slab * ncol + column; slab s covers z in [s*zs, (s+1)*zs)
    const int nslab = MDSF_OWNERS / ncol;
    const int zs = (nz + nslab - 1) / nslab;
    const int mycol = pt % ncol, myslab = pt / ncol;
    const int mycx = mycol / gp.ty, mycy = mycol % gp.ty;
    const int zlo = myslab * zs, zhi = min(zlo + zs, nz);
    const bool owner_valid = X0 + mycx < gp.n[0] && Y0 + mycy < gp.n[1] && zlo < zhi;
    double* col = mytile + (size_t)mycol * nzp;
    const int colbits = ncol;                                  // owner words: 32/ncol slabs per 32-bit word
    const int slabs_per_word = 32 / colbits;

    for (unsigned cb = lbeg; cb < lend; cb += chunk) {
        const int npair = (int)min((unsigned)chunk, lend - cb);
        // ---------------- A0: clip one pair per thread
        unsigned colmask = 0, slabmask = 0; int nxy = 0, nzc = 0;
        PairSlot s;
        if (pt < npair) {
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
            const int nzr = 2 * Az;
            const int kA = min(max(-s.pz0, 0), nzr), kB = min(max(nz - s.pz0, 0), nzr);
            s.nz = (short)nzr; s.kA = (short)kA; s.kB = (short)kB; s.pad0 = 0;
            // fold (dens.py:95-107): padding cells move by -+N_z, except in the 8 corner regions
            // whose z block is chosen by the y side: there the shift is +-Nborder when sy != sz
            const bool corner = (sx != 0 && sy != 0 && gp.fold_mode == 0);
            s.shlo = (corner && sy != -1) ? gp.nb : nz;
            s.shhi = (corner && sy != 1) ? -gp.nb : -nz;
            nxy = w * h; nzc = (nxy > 0) ? nzr : 0;
            if (nxy > 0) {
                for (int cx = cx0; cx < cx0 + w; ++cx)
                    colmask |= (((h >= 32) ? 0xffffffffu : ((1u << h) - 1u)) << (cx * gp.ty + cy0));
                if (kA > 0) { const int a0 = s.pz0 + s.shlo, a1 = s.pz0 + kA - 1 + s.shlo;
                              for (int sl = a0 / zs; sl <= a1 / zs; ++sl) slabmask |= 1u << sl; }
                if (kB > kA) { const int a0 = s.pz0 + kA, a1 = s.pz0 + kB - 1;
                               for (int sl = a0 / zs; sl <= a1 / zs; ++sl) slabmask |= 1u << sl; }
                if (nzr > kB) { const int a0 = s.pz0 + kB + s.shhi, a1 = s.pz0 + nzr - 1 + s.shhi;
                                for (int sl = a0 / zs; sl <= a1 / zs; ++sl) slabmask |= 1u << sl; }
            }
        }
        // per-part exclusive scan of (nxy, nzc) over the 128 threads
        int ixy = nxy, iz = nzc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t1 = __shfl_up_sync(0xffffffffu, ixy, d), t2 = __shfl_up_sync(0xffffffffu, iz, d);
            if (lane >= d) { ixy += t1; iz += t2; }
        }
        if (lane == 31) { scan_tmp[pw * 2] = ixy; scan_tmp[pw * 2 + 1] = iz; }
        // ballot transpose of the 32 x 128 (pair x owner) hit matrix of this warp:
        // hitT[warp][o] = pairs (bit = lane) that touch owner o, list order = bit order
        for (int word = 0; word < MDSF_OWNERS / 32; ++word) {
            unsigned ow = 0;                                   // this pair's hits on owners [32*word, 32*word+32)
            for (int j = 0; j < slabs_per_word; ++j)
                if ((slabmask >> (word * slabs_per_word + j)) & 1u) ow |= colmask << (j * colbits);
            unsigned mine = 0;
#pragma unroll 8
            for (int c = 0; c < 32; ++c) {
                const unsigned b = __ballot_sync(0xffffffffu, (ow >> c) & 1u);
                if (lane == c) mine = b;
            }
            hitT[pw * MDSF_OWNERS + word * 32 + lane] = mine;
        }
        part_barrier(part);
        int bxy = 0, bz = 0;
        for (int wv = 0; wv < pw; ++wv) { bxy += scan_tmp[wv * 2]; bz += scan_tmp[wv * 2 + 1]; }
        if (pt < npair) {
            s.offxy = bxy + ixy - nxy; s.offz = bz + iz - nzc;
            slots[pt] = s;
        }
        if (pt == 127) { scan_tmp[8] = bxy + ixy; scan_tmp[9] = bz + iz; }
        part_barrier(part);
        const int totxy = scan_tmp[8], totz = scan_tmp[9];

        // ---------------- A1: dense evaluation of the Gaussian factor tables
        if (gp.separable) {
            for (int e = pt; e < totxy + totz; e += 128) {
                const bool isz = e >= totxy;
                const int ee = isz ? e - totxy : e;
                int lo = 0, hi = npair - 1;       // last slot whose offset <= ee
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    const int o = isz ? slots[mid].offz : slots[mid].offxy;
                    if (o <= ee) lo = mid; else hi = mid - 1;
                }
                const PairSlot& p = slots[lo];
                const double t2 = tt.two_sig2[p.type];
                if (!isz) {
                    const int hh = (int)(p.rect >> 24);
                    const int li = ee - p.offxy;
                    const int lx = li / hh, ly = li - lx * hh;
                    // b = r - (i - B)*dr with the product rounded on its own (dens.py:252-256,299)
                    const double bx = __dsub_rn(p.rx, __dmul_rn((double)(p.px0 + lx), gp.dr[0]));
                    const double by = __dsub_rn(p.ry, __dmul_rn((double)(p.py0 + ly), gp.dr[1]));
                    const double c0 = gp.u[0] * bx + gp.u[3] * by;     // sum_m ucell[m][0] b_m (dens.py:301)
                    const double c1 = gp.u[1] * bx + gp.u[4] * by;
                    tblxy[ee] = exp(-(c0 * c0 + c1 * c1) / t2);
                } else {
                    const int k = ee - p.offz;
                    const double bzv = __dsub_rn(p.rz, __dmul_rn((double)(p.pz0 + k), gp.dr[2]));
                    const double c2 = gp.u[8] * bzv;
                    tblz[ee] = tt.amp[p.type] * exp(-(c2 * c2) / t2);
                }
            }
            part_barrier(part);
        }

        // ---------------- B: every owner adds its hits, in list order, into cells only it writes
        if (owner_valid) {
            const int nwarp_used = (npair + 31) >> 5;
            for (int wv = 0; wv < nwarp_used; ++wv) {
                unsigned m = hitT[wv * MDSF_OWNERS + pt];
                while (m) {
                    const int i = wv * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const PairSlot& p = slots[i];
                    const int cx0 = (int)(p.rect & 0xff), cy0 = (int)((p.rect >> 16) & 0xff), hh = (int)(p.rect >> 24);
                    const int lx = mycx - cx0, ly = mycy - cy0;
                    const int pz0 = p.pz0, kA = p.kA, kB = p.kB, nzr = p.nz;
                    if (gp.separable) {
                        const double exy = tblxy[p.offxy + lx * hh + ly];
                        const double* ez = tblz + p.offz;
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
    if (FUSE_ZFFT) fft_tile_z(tile_re, tile_im, twr, twi, zplan, ncol, nzp, gp.pad_shift);
    for (int c = 0; c < ncol; ++c) {
        const int x = X0 + c / gp.ty, y = Y0 + c % gp.ty;
        if (x < gp.n[0] && y < gp.n[1]) {
            double2* dst = vol + (((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz;
            for (int z = threadIdx.x; z < nz; z += blockDim.x) {
                const int a = c * nzp + z + (z >> gp.pad_shift);
                dst[z] = make_double2(tile_re[a], tile_im[a]);
            }
        }
    }
}
