// K3: deterministic owner-computes Gaussian splat of a column tile, fused with the z-axis FFT.
//
// Replaces the per-atom Python loop dens.py:283-308 and the 26-region fold dens.py:86-108.
// A CTA owns tx*ty (x,y) columns over all z for ONE pair of frames (frame 2q -> real part,
// frame 2q+1 -> imaginary part; threads 0..127 serve the real part, 128..255 the imaginary
// part).  Its sorted pair list (K2) is consumed in chunks:
//   A0  one thread per pair: clip the atom image against the tile, build the column hit mask;
//       a ballot transpose turns the per-pair masks into per-column ordered hit lists;
//   A1  the Gaussian factors are evaluated densely, one table entry per thread:
//       exy[pair][column] = exp(-(c0^2+c1^2)/(2 sigma^2)),  ez[pair][k] = Nel/sigma^3 exp(-c2^2/(2 sigma^2));
//   B   every column is owned by one lane group which walks ITS hits in list order and adds
//       exy*ez[k] into its shared-memory column: no atomics, fixed summation order, so the
//       density is bitwise reproducible.  The fold (incl. the corner rule of dens.py:107) is an
//       index map applied on the fly; no padded array exists.
// Afterwards the tile is either stored (debug / library-FFT path) or transformed along z in
// place and stored in position space (see mdsf_fft.cuh).
#pragma once
#include "mdsf_common.cuh"
#include "mdsf_fft.cuh"

struct PairSlot {
    double rx, ry, rz;     // atom coordinate (float64 value of the coords dtype)
    int px0, py0, pz0;     // padded-grid index of the first clipped column / first z cell
    int type;
    unsigned rect;         // cx0 | w << 8 | cy0 << 16 | h << 24   (tile-relative clip rectangle)
    int nz;                // 2*Az
    int flags;             // bit0: image is outside the cell in x AND y; bits1-2: sy+1
    int offxy, offz;       // table offsets
    int pad_;
};

__device__ __forceinline__ void part_barrier(int part) {
    asm volatile("bar.sync %0, %1;" ::"r"(part + 1), "r"(128) : "memory");
}

template <bool FUSE_ZFFT>
__global__ void __launch_bounds__(256)
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

    // ---- shared memory carve-up
    double* tile_re = smem;                                   // [ncol][nzp]  frame 2q
    double* tile_im = tile_re + (size_t)ncol * nzp;           // [ncol][nzp]  frame 2q+1
    double* twr = tile_im + (size_t)ncol * nzp;               // [nz] z twiddles (fused FFT only)
    double* twi = twr + (FUSE_ZFFT ? gp.n[2] : 0);
    double* tables = twi + (FUSE_ZFFT ? gp.n[2] : 0);
    const size_t tbl_per_part = (size_t)chunk * (xycap + zcap);
    double* tblxy = tables + part * tbl_per_part;             // [chunk*xycap]
    double* tblz = tblxy + (size_t)chunk * xycap;             // [chunk*zcap]
    PairSlot* slots_all = reinterpret_cast<PairSlot*>(tables + 2 * tbl_per_part);
    unsigned* hit_all = reinterpret_cast<unsigned*>(slots_all + 2 * chunk);
    int* scan_all = reinterpret_cast<int*>(hit_all + 2 * 4 * 32);
    PairSlot* slots = slots_all + part * chunk;               // [chunk]
    unsigned* hitT = hit_all + part * 4 * 32;                 // [4 warps][32 columns]
    int* scan_tmp = scan_all + part * 16;                     // warp totals + chunk totals

    double* mytile = part ? tile_im : tile_re;
    for (int i = threadIdx.x; i < 2 * ncol * nzp; i += blockDim.x) tile_re[i] = 0.0;
    if (FUSE_ZFFT) load_twiddles(twr, twi, twz, gp.n[2]);
    __syncthreads();

    const int f = 2 * q + part;
    const unsigned lbeg = tile_start[f * ntiles + tile], lend = tile_start[f * ntiles + tile + 1];

    // column ownership for phase B
    int G = 128 / ncol; G = G < 4 ? 4 : (G > 32 ? 32 : G);
    const int mycol = pt / G, klane = pt % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    const int mycx = mycol / gp.ty, mycy = mycol % gp.ty;
    const bool colvalid = mycol < ncol && X0 + mycx < gp.n[0] && Y0 + mycy < gp.n[1];

    for (unsigned cb = lbeg; cb < lend; cb += chunk) {
        const int npair = (int)min((unsigned)chunk, lend - cb);
        // ---------------- A0: clip one pair per thread
        unsigned mask = 0; int nxy = 0, nzc = 0;
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
            int cx0 = max(xlo - sx * gp.n[0] - X0, 0), cx1 = min(xhi - sx * gp.n[0] - X0, gp.tx);
            int cy0 = max(ylo - sy * gp.n[1] - Y0, 0), cy1 = min(yhi - sy * gp.n[1] - Y0, gp.ty);
            const int w = max(cx1 - cx0, 0), h = max(cy1 - cy0, 0);
            s.rx = rec.r[0]; s.ry = rec.r[1]; s.rz = rec.r[2];
            s.px0 = X0 + cx0 + sx * gp.n[0];
            s.py0 = Y0 + cy0 + sy * gp.n[1];
            s.pz0 = rec.ir[2] - Az;
            s.type = rec.type;
            s.rect = (unsigned)cx0 | ((unsigned)w << 8) | ((unsigned)cy0 << 16) | ((unsigned)h << 24);
            s.nz = 2 * Az;
            s.flags = ((sx != 0 && sy != 0) ? 1 : 0) | ((sy + 1) << 1);
            nxy = w * h; nzc = (nxy > 0) ? 2 * Az : 0;
            for (int cx = cx0; cx < cx0 + w; ++cx)
                mask |= (((h >= 32) ? 0xffffffffu : ((1u << h) - 1u)) << (cx * gp.ty + cy0));
        }
        // per-part exclusive scan of (nxy, nzc) over the 128 threads
        int ixy = nxy, iz = nzc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t1 = __shfl_up_sync(0xffffffffu, ixy, d), t2 = __shfl_up_sync(0xffffffffu, iz, d);
            if (lane >= d) { ixy += t1; iz += t2; }
        }
        if (lane == 31) { scan_tmp[pw * 2] = ixy; scan_tmp[pw * 2 + 1] = iz; }
        // ballot transpose: hitT[warp][c] = pairs of this warp whose image covers column c
        unsigned mine = 0;
        for (int c = 0; c < ncol; ++c) {
            const unsigned b = __ballot_sync(0xffffffffu, (mask >> c) & 1u);
            if (lane == c) mine = b;
        }
        hitT[pw * 32 + lane] = mine;
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

        // ---------------- B: each column group adds its hits, in list order
        if (colvalid) {
            const int nwarp_used = (npair + 31) >> 5;
            for (int wv = 0; wv < nwarp_used; ++wv) {
                unsigned m = hitT[wv * 32 + mycol];
                while (m) {
                    const int i = wv * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const PairSlot& p = slots[i];
                    const int cx0 = (int)(p.rect & 0xff), cy0 = (int)((p.rect >> 16) & 0xff), hh = (int)(p.rect >> 24);
                    const int lx = mycx - cx0, ly = mycy - cy0;
                    const bool corner = p.flags & 1;
                    const int sy = ((p.flags >> 1) & 3) - 1;
                    double* col = mytile + (size_t)mycol * nzp;
                    // z cells by side (low padding, cell, high padding): within one side the
                    // destinations are distinct, so the lanes of the group never collide
                    const int kA = min(max(-p.pz0, 0), p.nz), kB = min(max(gp.n[2] - p.pz0, 0), p.nz);
                    if (gp.separable) {
                        const double exy = tblxy[p.offxy + lx * hh + ly];
                        const double* ez = tblz + p.offz;
#pragma unroll 1
                        for (int seg = 0; seg < 3; ++seg) {
                            const int k0 = seg == 0 ? 0 : (seg == 1 ? kA : kB), k1 = seg == 0 ? kA : (seg == 1 ? kB : p.nz);
                            if (k1 > k0) {
                                for (int k = k0 + klane; k < k1; k += G) {
                                    const int cz = fold_z(p.pz0 + k, gp.n[2], gp.nb, corner, sy, gp.fold_mode);
                                    col[cz + (cz >> gp.pad_shift)] += exy * ez[k];
                                }
                                __syncwarp(gmask);
                            }
                        }
                    } else {
                        // general ucell: one exp per cell, exactly the reference's expression
                        const double bx = __dsub_rn(p.rx, __dmul_rn((double)(p.px0 + lx), gp.dr[0]));
                        const double by = __dsub_rn(p.ry, __dmul_rn((double)(p.py0 + ly), gp.dr[1]));
                        const double t2 = tt.two_sig2[p.type], amp = tt.amp[p.type];
#pragma unroll 1
                        for (int seg = 0; seg < 3; ++seg) {
                            const int k0 = seg == 0 ? 0 : (seg == 1 ? kA : kB), k1 = seg == 0 ? kA : (seg == 1 ? kB : p.nz);
                            if (k1 > k0) {
                                for (int k = k0 + klane; k < k1; k += G) {
                                    const double bzv = __dsub_rn(p.rz, __dmul_rn((double)(p.pz0 + k), gp.dr[2]));
                                    const double c0 = gp.u[0] * bx + gp.u[3] * by + gp.u[6] * bzv;
                                    const double c1 = gp.u[1] * bx + gp.u[4] * by + gp.u[7] * bzv;
                                    const double c2 = gp.u[2] * bx + gp.u[5] * by + gp.u[8] * bzv;
                                    const int cz = fold_z(p.pz0 + k, gp.n[2], gp.nb, corner, sy, gp.fold_mode);
                                    col[cz + (cz >> gp.pad_shift)] += amp * exp(-(c0 * c0 + c1 * c1 + c2 * c2) / t2);
                                }
                                __syncwarp(gmask);
                            }
                        }
                    }
                }
            }
        }
        part_barrier(part);
    }
    __syncthreads();

    const int nz = gp.n[2];
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
    if (FUSE_ZFFT) fft_tile<true>(tile_re, tile_im, twr, twi, zplan, ncol, nzp, 1, gp.pad_shift);
    for (int i = threadIdx.x; i < ncol * nz; i += blockDim.x) {
        const int c = i / nz, z = i - c * nz;
        const int x = X0 + c / gp.ty, y = Y0 + c % gp.ty;
        if (x < gp.n[0] && y < gp.n[1]) {
            const int a = c * nzp + z + (z >> gp.pad_shift);
            vol[(((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz + z] = make_double2(tile_re[a], tile_im[a]);
        }
    }
}
