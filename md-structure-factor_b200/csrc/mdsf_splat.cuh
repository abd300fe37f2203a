// K3: deterministic register-tiled Gaussian splat of a column tile, fused with the z-axis FFT.
//
// Replaces the per-atom Python loop dens.py:283-308 and the 26-region fold dens.py:86-108.
// A CTA owns 2^LCOL (x,y) columns over all z for ONE pair of frames (frame 2q -> real part, frame 2q+1 ->
// imaginary part).  A warp owns one z slab of the tile at a time: lane = (column, z lane), 8 consecutive cells per
// lane, held in REGISTERS as 64-bit fixed-point sums (LSB = 2^-52 of the largest Nel/sigma^3).  The warp walks its
// own (tile, slab) list of pre-clipped pair records (K2): per record the slab's window of the atom's EZ table and
// the tile's slices of its EX / EY (/ cross-term) tables arrive in a small per-warp staging ring by cp.async, NS-1
// records ahead; entries outside the clip box are zero-filled, so the accumulation is branch-free:
//     acc[k] += int(EX[cx] * EY[cy] * C[c] * EZ[zl*8 + k])          one DFMA (magic-number rounding) + one 64-bit add
// Integer addition commutes, so the density is bitwise reproducible whatever order K2's atomics filled the lists
// in -- no float atomics, no shared-memory atomics, no CTA barriers inside the splat.  The fold (incl. the corner
// rule of dens.py:107) was resolved by K2: every record is one box in destination space.
// Afterwards the slab sums are converted to fp64 into the shared-memory tile, which is transformed along z in
// place (native FFT path) and stored: the density never touches HBM.
#pragma once
#include "mdsf_common.cuh"
#include "mdsf_fft.cuh"

#define MDSF_SPLAT_THREADS (32 * MDSF_SPLAT_WARPS)
#define MDSF_MAGIC 4503599627370496.0              // 2^52: fma(a, b, 2^52) = 2^52 + rint(a*b) for 0 <= a*b < 2^52
#define MDSF_MAGIC_BITS 0x4330000000000000LL

enum { SPLAT_ORTHO = 0,      // separable ucell, no in-plane cross term
       SPLAT_MONO = 1,       // separable ucell with cross-term table (monoclinic, theta != 90)
       SPLAT_GENERAL = 2,    // arbitrary 3x3 ucell: one exp per cell, exactly the reference's expression
       SPLAT_DENSITY = 3 };  // no atoms: real densities d1[frame][x][y][z] are the source (RANDOM_NOISE mode, dens.py:279-280)

template <int LCOL> struct SplatGeom {
    static constexpr int NCOL = 1 << LCOL, LTY = LCOL / 2, TX = 1 << ((LCOL + 1) / 2), TY = 1 << LTY;
    static constexpr int ZL = 32 >> LCOL, ZW = ZL * 8;              // z lanes per warp, slab width
    static constexpr int NS = (LCOL == 2) ? 2 : 3;                  // staging slots per warp
    static constexpr int SLOT = (ZW + TX + TY + NCOL + 1) & ~1;     // doubles: [EZ: ZW][EX: TX][EY: TY][C: NCOL]
    static constexpr int WARP_BYTES = 32 * 16 + 32 * 8 + NS * SLOT * 8;
};

__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int LCOL, int MODE>
__global__ void __launch_bounds__(MDSF_SPLAT_THREADS, 2)
splat_zfft_kernel(const PairRec* __restrict__ prec, const PairAux* __restrict__ paux, const unsigned* __restrict__ start,
                  const AtomRec* __restrict__ recs, const double* __restrict__ tables,
                  const double* __restrict__ src_density, int nframes,
                  double2* __restrict__ vol, double2* __restrict__ dens_dump, GridParams gp, TypeTable tt, FftPlan zplan,
                  const double2* __restrict__ twz, int* __restrict__ err_flag, int FUSE /* z FFT fused (native path) */)
{
    using G = SplatGeom<LCOL>;
    constexpr int NCOL = G::NCOL, TX = G::TX, TY = G::TY, LTY = G::LTY, ZW = G::ZW, NS = G::NS, SLOT = G::SLOT;
    extern __shared__ double smem[];
    const int nzp = gp.nzp, nz = gp.n[2];
    double* tile_re = smem;                                   // [NCOL][nzp]  frame 2q
    double* tile_im = tile_re + (size_t)NCOL * nzp;           // [NCOL][nzp]  frame 2q+1
    char* area = reinterpret_cast<char*>(tile_im + (size_t)NCOL * nzp);     // per-warp staging; z twiddles afterwards
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, q = blockIdx.y;
    const int X0 = (tile / gp.nty) * TX, Y0 = (tile % gp.nty) * TY;
    const int c = lane & (NCOL - 1), zl = lane >> LCOL;       // my column of the tile, my z lane of the slab
    const int cx = c >> LTY, cy = c & (TY - 1);
    const int ntiles = gp.ntx * gp.nty;
    const bool col_ok = X0 + cx < gp.n[0] && Y0 + cy < gp.n[1];

    PairRec* rbuf = reinterpret_cast<PairRec*>(area + (size_t)warp * G::WARP_BYTES);
    PairAux* abuf = reinterpret_cast<PairAux*>(reinterpret_cast<char*>(rbuf) + 32 * 16);
    double* slots = reinterpret_cast<double*>(reinterpret_cast<char*>(rbuf) + 32 * 16 + 32 * 8);
    bool ovf = false;

    for (int s = warp; s < gp.nslab; s += MDSF_SPLAT_WARPS) {
        const int zbase = s * ZW + zl * 8;                    // my 8 cells: z = zbase + k
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const int f = 2 * q + part;
            double* col = (part ? tile_im : tile_re) + (size_t)c * nzp;
            if (MODE == SPLAT_DENSITY) {
                const bool live = f < nframes && col_ok;
                const double* src = src_density + (((long long)f * gp.n[0] + (X0 + cx)) * gp.n[1] + (Y0 + cy)) * nz;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int z = zbase + k;
                    if (z < nz) col[z + (z >> gp.pad_shift)] = live ? src[z] : 0.0;
                }
                continue;
            }
            long long acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0;
            const unsigned key = (unsigned)(f * ntiles + tile) * (unsigned)gp.nslab + (unsigned)s;
            const unsigned lbeg = start[key], lend = start[key + 1];
            const int n = (int)(lend - lbeg);
            PairRec nxt = make_uint4(0u, 0u, 0u, 0u);
            PairAux nxa = make_uint2(0u, 0u);
            if (lane < n) { nxt = prec[lbeg + lane]; if (MODE != SPLAT_ORTHO) nxa = paux[lbeg + lane]; }
            for (int b = 0; b < n; b += 32) {
                const int m = min(32, n - b);
                __syncwarp();                                  // the previous batch is consumed
                rbuf[lane] = nxt;
                if (MODE != SPLAT_ORTHO) abuf[lane] = nxa;
                __syncwarp();
                if (b + 32 + lane < n) { nxt = prec[lbeg + b + 32 + lane]; if (MODE != SPLAT_ORTHO) nxa = paux[lbeg + b + 32 + lane]; }
                if (MODE == SPLAT_GENERAL) {
                    // one exp per cell: amp * exp(-|b . ucell|^2 / (2 sigma^2))  (dens.py:299-308)
                    for (int i = 0; i < m; ++i) {
                        const PairRec r = rbuf[i];
                        const unsigned g = r.w;
                        const int cx0 = g & 7, cx1 = (g >> 3) & 15, cy0 = (g >> 7) & 7, cy1 = (g >> 10) & 15;
                        const int zoff = (g >> 14) & 127, zend = (g >> 21) & 127;
                        if (cx < cx0 || cx >= cx1 || cy < cy0 || cy >= cy1) continue;
                        const AtomRec* ar = recs + (long long)f * gp.natoms + abuf[i].x;
                        const double rx = ar->r[0], ry = ar->r[1], rz = ar->r[2];
                        const int type = ar->type;
                        const double bx = __dsub_rn(rx, __dmul_rn((double)((int)r.y + cx), gp.dr[0]));
                        const double by = __dsub_rn(ry, __dmul_rn((double)((int)r.z + cy), gp.dr[1]));
                        const double t2 = tt.two_sig2[type], amp = tt.amp[type];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int zz = zl * 8 + k;
                            if (zz < zoff || zz >= zend) continue;
                            const double bzv = __dsub_rn(rz, __dmul_rn((double)((int)r.x + zz), gp.dr[2]));
                            const double c0 = gp.u[0] * bx + gp.u[3] * by + gp.u[6] * bzv;
                            const double c1 = gp.u[1] * bx + gp.u[4] * by + gp.u[7] * bzv;
                            const double c2 = gp.u[2] * bx + gp.u[5] * by + gp.u[8] * bzv;
                            acc[k] += __double2ll_rn(amp * exp(-(c0 * c0 + c1 * c1 + c2 * c2) / t2) * gp.fx_scale);
                        }
                    }
                    continue;
                }
                // ---- separable ucell: cp.async staging ring, NS-1 records ahead
                constexpr int E = ZW + TX + TY + (MODE == SPLAT_MONO ? NCOL : 0);
                constexpr int ROUNDS = (E + 31) / 32;
                auto produce = [&](int i, int slot) {
                    const PairRec r = rbuf[i];
                    const unsigned g = r.w;
                    const int cx0 = g & 7, cx1 = (g >> 3) & 15, cy0 = (g >> 7) & 7, cy1 = (g >> 10) & 15;
                    const int zoff = (g >> 14) & 127, zend = (g >> 21) & 127;
                    double* dst = slots + slot * SLOT;
#pragma unroll
                    for (int rd = 0; rd < ROUNDS; ++rd) {
                        const int e = lane + 32 * rd;
                        if (e >= E) break;
                        bool valid;
                        const double* src;
                        if (e < ZW) { valid = e >= zoff && e < zend; src = tables + ((int)r.x + e); }
                        else if (e < ZW + TX) { const int xe = e - ZW; valid = xe >= cx0 && xe < cx1; src = tables + ((int)r.y + xe); }
                        else if (e < ZW + TX + TY) { const int ye = e - ZW - TX; valid = ye >= cy0 && ye < cy1; src = tables + ((int)r.z + ye); }
                        else {
                            const int ce = e - ZW - TX - TY, xe = ce >> LTY, ye = ce & (TY - 1);
                            valid = xe >= cx0 && xe < cx1 && ye >= cy0 && ye < cy1;
                            const PairAux a = abuf[i];
                            src = tt.ctab + ((int)a.x + xe * (int)a.y + ye);
                        }
                        if (!valid) src = tables;
                        cp_async8_zfill(dst + e, src, valid);
                    }
                };
                int ps = 0, cs = 0;
#pragma unroll
                for (int i = 0; i < NS - 1; ++i) {
                    if (i < m) produce(i, ps);
                    cp_async_commit();
                    ps = (ps + 1 == NS) ? 0 : ps + 1;
                }
                for (int i = 0; i < m; ++i) {
                    if (i + NS - 1 < m) produce(i + NS - 1, ps);
                    cp_async_commit();
                    ps = (ps + 1 == NS) ? 0 : ps + 1;
                    cp_async_wait<NS - 1>();
                    __syncwarp();
                    const double* sl = slots + cs * SLOT;
                    double exy = sl[ZW + cx] * sl[ZW + TX + cy];
                    if (MODE == SPLAT_MONO) exy *= sl[ZW + TX + TY + c];
                    const double2* ez2 = reinterpret_cast<const double2*>(sl + zl * 8);
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        const double2 e = ez2[k2];
                        acc[2 * k2]     += __double_as_longlong(__fma_rn(exy, e.x, MDSF_MAGIC)) - MDSF_MAGIC_BITS;
                        acc[2 * k2 + 1] += __double_as_longlong(__fma_rn(exy, e.y, MDSF_MAGIC)) - MDSF_MAGIC_BITS;
                    }
                    cs = (cs + 1 == NS) ? 0 : cs + 1;
                    __syncwarp();                              // slot cs-1 may be refilled
                }
            }
            // fixed point -> fp64 (overflow: a cell held > 2048 peak amplitudes)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int z = zbase + k;
                ovf |= (acc[k] < 0) | (acc[k] >= (1LL << 62));
                if (z < nz) col[z + (z >> gp.pad_shift)] = (double)acc[k] * gp.fx_inv;
            }
        }
    }
    if (ovf) atomicExch(err_flag, 2);
    cp_async_wait<0>();
    __syncthreads();

    double* twr = reinterpret_cast<double*>(area);            // z twiddles reuse the staging area
    double* twi = twr + nz;
    if (FUSE) load_twiddles(twr, twi, twz, nz);
    if (dens_dump != nullptr) {                               // parity tap: plain [q][x][y][z] pairs, before any FFT
        for (int i = threadIdx.x; i < NCOL * nz; i += blockDim.x) {
            const int cc = i / nz, z = i - cc * nz;
            const int x = X0 + (cc >> LTY), y = Y0 + (cc & (TY - 1));
            if (x < gp.n[0] && y < gp.n[1]) {
                const int a = cc * nzp + z + (z >> gp.pad_shift);
                dens_dump[(((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz + z] = make_double2(tile_re[a], tile_im[a]);
            }
        }
    }
    __syncthreads();

    const long long ncell = (long long)gp.n[0] * gp.n[1] * nz;
    const long long cs_ = (long long)gp.n[0] * gp.n[1] * gp.lw;      // chunk stride
    double2* volq = vol + (long long)q * ncell;
    if (FUSE) fft_tile_z(tile_re, tile_im, twr, twi, zplan, NCOL, nzp, gp.pad_shift);      // ends with a barrier
    // store: runs of TY * lw contiguous cells (chunked layout) / whole columns (plain layout)
    if (gp.lw == nz) {
        for (int cc = 0; cc < NCOL; ++cc) {
            const int x = X0 + (cc >> LTY), y = Y0 + (cc & (TY - 1));
            if (x >= gp.n[0] || y >= gp.n[1]) continue;
            double2* dst = volq + ((long long)x * gp.n[1] + y) * nz;
            const double* sre = tile_re + (size_t)cc * nzp;
            const double* sim = tile_im + (size_t)cc * nzp;
            for (int z = threadIdx.x; z < nz; z += blockDim.x) {
                const int a = z + (z >> gp.pad_shift);
                dst[z] = make_double2(sre[a], sim[a]);
            }
        }
    } else {
        const int llw = __ffs(gp.lw) - 1;
        const int per_x = TY * nz;                             // elements of one tile row x: [ch][cy][zw]
        for (int xx = 0; xx < TX; ++xx) {
            const int x = X0 + xx;
            if (x >= gp.n[0]) break;
            for (int j = threadIdx.x; j < per_x; j += blockDim.x) {
                const int ch = j >> (LTY + llw), yy = (j >> llw) & (TY - 1), zw = j & (gp.lw - 1);
                const int y = Y0 + yy;
                if (y >= gp.n[1]) continue;
                const int z = (ch << llw) + zw;
                const int a = ((xx << LTY) + yy) * nzp + z + (z >> gp.pad_shift);
                volq[(long long)ch * cs_ + ((long long)x * gp.n[1] + y) * gp.lw + zw] = make_double2(tile_re[a], tile_im[a]);
            }
        }
    }
}
