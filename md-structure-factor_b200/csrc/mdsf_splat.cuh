// K3: deterministic register-tiled Gaussian splat of a column tile, fused with the z-axis FFT.
//
// Replaces the per-atom Python loop dens.py:283-308 and the 26-region fold dens.py:86-108.
// A CTA owns 2^LCOL (x,y) columns over all z for ONE pair of frames (frame 2q -> real part, frame 2q+1 ->
// imaginary part).  A warp owns one z slab of the tile at a time and holds its cells in REGISTERS as 64-bit fixed-point
// sums (LSB = 2^-52 of the largest Nel/sigma^3).  It walks its own (tile, slab) list of pair records (K2); per record:
//     acc[col][k] += int(EX[i] * EY[j] * C[j][i] * EZ[z])   one DFMA (magic-number rounding) + one 64-bit add
// EZ entries outside the record's z window are zero, EX / EY entries of columns outside the stamp come from the zero
// pads of the atom's table block, so the accumulation is branch-free.  Integer addition commutes, so the density is
// bitwise reproducible whatever order K2's atomics filled the lists in -- no float atomics, no shared-memory atomics,
// no CTA barriers inside the splat.  The fold (incl. the corner rule of dens.py:107) was resolved by K2: every record
// is one box in destination space.  Two lane mappings:
//   * z-lane (ZL, see ZLaneGeom): lanes walk z, a lane holds up to 8 columns; products of a batch of records formed
//     up front, EZ straight from global memory two records ahead;
//   * half-warp lists (SUB = 2, 4x4-column tiles): lane = (tile row, z lane) x TX columns x KZ = 8/TX cells, two lists
//     side by side per warp, table slices staged by cp.async in a per-warp ring NS-1 records ahead.
// Afterwards the slab sums are converted to fp64 into the shared-memory tile, which is transformed along z in
// place (native FFT path; compile-time radix stages on an interleaved 16-byte tile for Nz = 64 / 256 / 512 / 768 /
// 1024, run-time stages on split re / im planes otherwise) and stored: the density never touches HBM.
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

// SUB = lists a warp walks side by side: 1 (the whole warp owns one slab) or 2 (each half-warp owns a slab of half
// the width: small stamps touch few cells of a slab, two records per instruction stream halve the issue and
// shared-memory cost per record at the price of ~15% more records)
template <int LCOL, int SUB> struct SplatGeom {
    static constexpr int NCOL = 1 << LCOL, LTY = LCOL / 2, TX = 1 << ((LCOL + 1) / 2), TY = 1 << LTY;
    static constexpr int LTX = (LCOL + 1) / 2;
    static constexpr int G = 32 / SUB;                              // lanes of one group
    static constexpr int KZ = 8 / TX, ZLN = G / TY, ZW = ZLN * KZ;  // cells per lane in z, z lanes per group, slab width
    static constexpr int NS = (LCOL == 2) ? 2 : 3;                  // staging slots per warp
    static constexpr int SLOTG = (ZW + TX + TY + NCOL + 1) & ~1;    // doubles per group: [EZ: ZW][EX: TX][EY: TY][C: TY][TX]
    static constexpr int SLOT = SUB * SLOTG;
    static constexpr int RB = 32 / (SUB * SUB);                     // records per group and batch (32 / 8)
    static constexpr int WARP_BYTES = SUB * RB * 16 + SUB * RB * 8 + NS * SLOT * 8;
};

// Z-lane variant (ZL): the whole warp owns one slab; a lane holds NCL = min(NCOL, 8) columns x KZ = 8 / NCL cells, lanes
// walk z (CG = NCOL / NCL column groups x ZLN = 32 / CG z lanes; slab width ZLN * KZ = 256 >> LCOL, the SUB = 1 slab).
// Per record a lane then needs its KZ entries of EZ -- loaded straight from the table block into registers, zero outside
// the record's z window, two records ahead -- and the NCL products EX*EY*C of its columns, which are the same for every
// z lane: the products of a whole batch of records are formed up front, one (record, column) pair per lane, and read
// back from shared memory as broadcast 16-byte loads.  The list bounds of all slabs of the CTA sit in shared memory,
// and a warp claims its next item and requests that item's first records before it works on the current one, so the
// only global round trip left on an item's critical path is the one of the factor tables.
#ifndef MDSF_ZL_RB
#define MDSF_ZL_RB 16
#endif
template <int LCOL> struct ZLaneGeom {
    static constexpr int NCOL = 1 << LCOL, NCL = NCOL < 8 ? NCOL : 8, CG = NCOL / NCL, ZLN = 32 / CG, KZ = 8 / NCL;
    static constexpr int RB = (256 / NCOL) < MDSF_ZL_RB ? (256 / NCOL) : MDSF_ZL_RB;   // records per batch
    static constexpr int RPI = 32 / NCOL;                                    // records whose products one warp instruction forms
    static constexpr int WARP_BYTES = RB * 16 + RB * 8 + RB * NCOL * 8;      // [RB] PairRec, [RB] PairAux, [RB][NCOL] products
};
#define MDSF_ZL_MAXSLAB 128      // list bounds of a CTA's slabs are staged in shared memory up to this many slabs

__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(unsigned smem_dst, const void* gsrc, bool valid) {
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// exact int64 -> fp64 (|v| < 2^62) from two 32-bit conversions and one FMA (the 64-bit conversion is a long sequence)
__device__ __forceinline__ double fx_to_double(long long v) {
    return __fma_rn(__int2double_rn((int)(v >> 32)), 4294967296.0, __uint2double_rn((unsigned)v));
}

// ------------------------------------------------------------------ z FFT of the tile, compile-time radices
// Position of cell z inside a column of the split re / im planes (run-time z stages): padded form p + (p >> pad).
__device__ __forceinline__ int zpos(int z, int pad) { return z + (z >> pad); }

// In-place decimation-in-frequency stages with every stride and index split a constant (same scheme and output order as
// fft_stage).  Only w = w_N^(n2 N/L) is read per butterfly (per-stage table tws[TOFF + n2], host-built); its powers
// w^2 .. w^(R-1) come from complex multiplications: the fp64 pipe idles at 15% here while the load/store pipe is the
// bottleneck, and seven strided twiddle loads per radix-8 butterfly were a third of a stage's shared-memory wavefronts.
// ---- interleaved tile (gp.zilv): [NCOL][nz + 1] complex cells of 16 bytes, frame 2q in .x, frame 2q+1 in .y.  The odd
// column stride puts the same cell of 8 consecutive columns into 8 different 16-byte bank groups, so with the column
// index fastest across the lanes every 16-byte access of a quarter-warp is conflict-free in every stage, whatever the
// row stride -- no swizzle arithmetic, half the shared-memory instructions of the split re / im planes.
__device__ __forceinline__ int tile_index(const GridParams& gp, int ncol, int c, int z, int part) {
    return gp.zilv ? (((c * gp.nzp + z) << 1) + part)
                   : (part * ncol * gp.nzp + c * gp.nzp + z + (z >> gp.pad_shift));
}

template <int NCOL, int N, int R, int L, int TOFF>
__device__ __forceinline__ void zstage2(double2* __restrict__ tile, const double2* __restrict__ tws, int nzp)
{
    constexpr int M = L / R, NBF = N / R;
    constexpr int ITEMS = NCOL * NBF;
    for (int it = threadIdx.x; it < ITEMS; it += MDSF_SPLAT_THREADS) {
        const int f = it % NCOL, bf = it / NCOL;            // constants: shifts / masks
        const int b = bf / M, n2 = bf % M;
        double2* col = tile + f * nzp + b * L + n2;
        double xr[R], xi[R];
#pragma unroll
        for (int j = 0; j < R; ++j) { const double2 v = col[j * M]; xr[j] = v.x; xi[j] = v.y; }
        Dft<R>::run(xr, xi, nullptr, nullptr, N);
        if (M > 1) {
            const double2 w1 = tws[TOFF + n2];
            double wr[R], wi[R];
            wr[1] = w1.x; wi[1] = w1.y;
#pragma unroll
            for (int k = 2; k < R; ++k) {                    // w^k = w^(k/2) * w^(k - k/2): depth log2(k)
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
#pragma unroll
        for (int k = 0; k < R; ++k) col[k * M] = make_double2(xr[k], xi[k]);
    }
    __syncthreads();
}
template <int NCOL> __device__ __forceinline__ void zfft2(double2* t, const double2* tws, int nzp, int nz) {
    if (nz == 64) {
        zstage2<NCOL, 64, 8, 64, 0>(t, tws, nzp); zstage2<NCOL, 64, 8, 8, 8>(t, tws, nzp);
    } else if (nz == 256) {
        zstage2<NCOL, 256, 4, 256, 0>(t, tws, nzp); zstage2<NCOL, 256, 8, 64, 64>(t, tws, nzp); zstage2<NCOL, 256, 8, 8, 72>(t, tws, nzp);
    } else if (nz == 512) {
        zstage2<NCOL, 512, 8, 512, 0>(t, tws, nzp); zstage2<NCOL, 512, 8, 64, 64>(t, tws, nzp); zstage2<NCOL, 512, 8, 8, 72>(t, tws, nzp);
    } else if (nz == 1024) {
        zstage2<NCOL, 1024, 4, 1024, 0>(t, tws, nzp); zstage2<NCOL, 1024, 4, 256, 256>(t, tws, nzp);
        zstage2<NCOL, 1024, 8, 64, 320>(t, tws, nzp); zstage2<NCOL, 1024, 8, 8, 328>(t, tws, nzp);
    } else {      // 768
        zstage2<NCOL, 768, 4, 768, 0>(t, tws, nzp); zstage2<NCOL, 768, 8, 192, 192>(t, tws, nzp);
        zstage2<NCOL, 768, 8, 24, 216>(t, tws, nzp); zstage2<NCOL, 768, 3, 3, 219>(t, tws, nzp);
    }
}

// z lengths with compile-time stages (and the interleaved tile); host and device agree through this function
__host__ __device__ inline bool zspec_length(int nz) { return nz == 64 || nz == 256 || nz == 512 || nz == 768 || nz == 1024; }

#ifndef MDSF_SPLAT_MINB
#define MDSF_SPLAT_MINB 2
#endif
template <int LCOL, int MODE, int SUB, bool ZL = false>
__global__ void __launch_bounds__(MDSF_SPLAT_THREADS, MDSF_SPLAT_MINB)
splat_zfft_kernel(const uint4* __restrict__ prec /* 32-byte pair records */, const unsigned* __restrict__ start,
                  const AtomRec* __restrict__ recs, const double* __restrict__ tables,
                  const double* __restrict__ src_density, int nframes,
                  double2* __restrict__ vol, double2* __restrict__ dens_dump, GridParams gp, TypeTable tt, FftPlan zplan,
                  const double2* __restrict__ twz, int* __restrict__ err_flag, int FUSE /* z FFT fused (native path) */,
                  const double2* __restrict__ tws_g /* per-stage twiddle tables (compile-time z path) */, int tws_n, int tws_off
                  /* > 0: byte offset of their own shared-memory region, filled by cp.async while the splat runs */)
{
    using G = SplatGeom<LCOL, SUB>;
    constexpr int NCOL = G::NCOL, TX = G::TX, TY = G::TY, LTY = G::LTY, LTX = G::LTX, ZW = G::ZW, NS = G::NS, SLOT = G::SLOT;
    constexpr int KZ = G::KZ, ZLN = G::ZLN, GL = G::G, SLOTG = G::SLOTG, RB = G::RB;
    extern __shared__ double smem[];
    __shared__ int s_next;                                    // next (slab group, part) item: warps claim them dynamically
    const int nzp = gp.nzp, nz = gp.n[2];
    double* tile_re = smem;                                   // [NCOL][nzp]  frame 2q
    double* tile_im = tile_re + (size_t)NCOL * nzp;           // [NCOL][nzp]  frame 2q+1
    char* area = reinterpret_cast<char*>(tile_im + (size_t)NCOL * nzp);     // per-warp staging; z twiddles afterwards
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, q = blockIdx.y;
    const int X0 = (tile / gp.nty) * TX, Y0 = (tile % gp.nty) * TY;
    const int grp = lane / GL, gl = lane % GL;                // my group (list) of the warp, my lane inside it
    const int zl = gl & (ZLN - 1), cy = gl / ZLN;             // my z lane of the slab, my row of the tile (columns c = i*TY + cy)
    const int ntiles = gp.ntx * gp.nty;

    constexpr int WB = ZL ? ZLaneGeom<LCOL>::WARP_BYTES : G::WARP_BYTES, NREC = ZL ? ZLaneGeom<LCOL>::RB : SUB * RB;
    PairRec* rbuf = reinterpret_cast<PairRec*>(area + (size_t)warp * WB);                       // [SUB][RB] / [32]
    PairAux* abuf = reinterpret_cast<PairAux*>(reinterpret_cast<char*>(rbuf) + NREC * 16);
    double* slots = reinterpret_cast<double*>(reinterpret_cast<char*>(rbuf) + NREC * 24);      // [NS][SUB][SLOTG] / [2][NCOL]
    const unsigned slots_s = (unsigned)__cvta_generic_to_shared(slots);
    bool ovf = false;
    const bool zspec = FUSE && tws_g != nullptr && gp.zilv;
    double2* tws_s = reinterpret_cast<double2*>(reinterpret_cast<char*>(smem) + tws_off);
    if (threadIdx.x == 0) s_next = MDSF_SPLAT_WARPS;
    if (zspec && tws_off > 0) {                               // the twiddles arrive behind the splat
        for (int i = threadIdx.x; i < tws_n; i += MDSF_SPLAT_THREADS) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(tws_s + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(tws_g + i) : "memory");
        }
        cp_async_commit();
    }
    __syncthreads();

    // staging roles of this lane: entry e = gl + GL * round of its group's slot [EZ: ZW][EX: TX][EY: TY][C: TY][TX].
    // small entry (e >= ZW): source index = base(kind) + soff + smul * (2 Ay);  kind 1: EX, 2: EY, 3: C
    constexpr int E = ZW + TX + TY + (MODE == SPLAT_MONO ? NCOL : 0);
    constexpr int ROUNDS = (E + GL - 1) / GL;
    int skind[ROUNDS], soff[ROUNDS], smul[ROUNDS];
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
        const int e = gl + GL * rd - ZW;
        skind[rd] = 0; soff[rd] = 0; smul[rd] = 0;
        if (e >= 0 && e < TX) { skind[rd] = 1; soff[rd] = e; }
        else if (e >= TX && e < TX + TY) { skind[rd] = 2; soff[rd] = e - TX; }
        else if (MODE == SPLAT_MONO && e >= TX + TY && e < TX + TY + NCOL) { skind[rd] = 3; soff[rd] = (e - TX - TY) >> LTX; smul[rd] = (e - TX - TY) & (TX - 1); }
    }

    if constexpr (ZL && (MODE == SPLAT_ORTHO || MODE == SPLAT_MONO)) {
        using Z = ZLaneGeom<LCOL>;
        constexpr int NCL = Z::NCL, ZLNZ = Z::ZLN, KZZ = Z::KZ, RBZ = Z::RB, RPI = Z::RPI;
        __shared__ unsigned s_bnd[2][MDSF_ZL_MAXSLAB + 1];                  // list begin of slab j of frame 2q / 2q+1 ([nslab]: end)
        const bool bnd_s = gp.nslab <= MDSF_ZL_MAXSLAB;
        const unsigned key0 = (unsigned)(2 * q * ntiles + tile) * (unsigned)gp.nslab, key1 = key0 + (unsigned)ntiles * (unsigned)gp.nslab;
        if (bnd_s) {
            for (int t = threadIdx.x; t < 2 * (gp.nslab + 1); t += MDSF_SPLAT_THREADS) {
                const int pt = t / (gp.nslab + 1), j = t - pt * (gp.nslab + 1);
                const unsigned key = (pt ? key1 : key0) + (unsigned)j;
                s_bnd[pt][j] = key ? start[key - 1] : 0u;
            }
            __syncthreads();
        }
        const int cg = lane / ZLNZ, zlz = lane % ZLNZ;                       // my column group, my z lane
        const int pc = lane & (NCOL - 1), prs = lane / NCOL;                 // product role: column, record of the instruction
        const int pi = pc >> LTY, pcy = pc & (TY - 1);                       // column pc = pi * TY + pcy
        double* pbuf = slots;                                                // [RBZ][NCOL]
        const int nitems = 2 * gp.nslab;
        auto bounds = [&](int it, unsigned& lb, int& cnt) {
            lb = 0u; cnt = 0;
            if (it >= nitems) return;
            const int sl = it >> 1, pt = it & 1;
            if (bnd_s) { lb = s_bnd[pt][sl]; cnt = (int)(s_bnd[pt][sl + 1] - lb); }
            else { const unsigned key = (pt ? key1 : key0) + (unsigned)sl; lb = key ? start[key - 1] : 0u; cnt = (int)(start[key] - lb); }
        };
        PairRec nxt = make_uint4(0u, 0u, 0u, 0u);
        PairAux nxa = make_uint2(0u, 0u);
        auto fetch = [&](unsigned lb, int cnt, int b) {                      // record b + lane of a list -> nxt / nxa
            if (lane < RBZ && b + lane < cnt) {
                nxt = prec[2 * (size_t)(lb + b + lane)];
                if (MODE != SPLAT_ORTHO) { const uint4 t = prec[2 * (size_t)(lb + b + lane) + 1]; nxa = make_uint2(t.x, t.y); }
            }
        };
        // Item order.  The list lengths of all items of the CTA are known up front (s_bnd), and the warps meet at a CTA barrier
        // before the z stages: the longest pair of lists decides when (ncu: 10 % of all warp time waited there).  With at most
        // two items per warp, warp 0 ranks the items by length (bitonic sort of 32 keys in registers) and warp w takes
        // the w-th longest and the w-th shortest.  Otherwise: static round robin (with one item claimed ahead a dynamic
        // counter hands the items out in start order anyway, and its atomic sat on every item's critical path).
        __shared__ unsigned char s_order[32];
        const bool ranked = bnd_s && nitems <= 32 && nitems > MDSF_SPLAT_WARPS && 2 * MDSF_SPLAT_WARPS >= nitems;
        if (ranked) {
            if (warp == 0) {
                unsigned lb0; int c0;
                bounds(lane, lb0, c0);
                unsigned key = lane < nitems ? ((unsigned)(0xffffff - min(c0, 0xffffff)) << 8) | (unsigned)lane : 0xffffffffu;
#pragma unroll
                for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        const unsigned o = __shfl_xor_sync(0xffffffffu, key, j);
                        const bool asc = (lane & k) == 0, low = (lane & j) == 0;
                        key = (asc == low) ? min(key, o) : max(key, o);
                    }
                s_order[lane] = (unsigned char)(key & 255u);         // position p: p-th longest list
            }
            __syncthreads();
        }
        auto item_at = [&](int k) -> int {                               // k-th item of this warp (>= nitems: none)
            if (!ranked) return warp + k * MDSF_SPLAT_WARPS;
            const int p = k == 0 ? warp : (k == 1 ? 2 * MDSF_SPLAT_WARPS - 1 - warp : nitems);
            return p < nitems ? (int)s_order[p] : nitems;
        };
        int turn = 0, item = item_at(0), n = 0;
        unsigned lbeg = 0u;
        bounds(item, lbeg, n);
        fetch(lbeg, n, 0);
        while (item < nitems) {
            const int item2 = item_at(++turn);
            unsigned lbeg2;
            int n2;
            bounds(item2, lbeg2, n2);
            const int s = item >> 1, part = item & 1;
            long long acc[NCL][KZZ];
#pragma unroll
            for (int c = 0; c < NCL; ++c)
#pragma unroll
                for (int k = 0; k < KZZ; ++k) acc[c][k] = 0;
            int b = 0;
            do {
                const int m = min(RBZ, n - b);
                __syncwarp();                                              // the previous batch is consumed
                if (lane < m) { rbuf[lane] = nxt; if (MODE != SPLAT_ORTHO) abuf[lane] = nxa; }
                __syncwarp();
                if (b + RBZ < n) fetch(lbeg, n, b + RBZ); else fetch(lbeg2, n2, 0);
                if (m > 0) {
                    auto ldez = [&](int k, double (&ez)[KZZ]) {
#pragma unroll
                        for (int kk = 0; kk < KZZ; ++kk) ez[kk] = 0.0;
                        if (k < m) {
                            const PairRec r = rbuf[k];
                            const unsigned zoff = r.w & 127u, zlen = r.w >> 7;
#pragma unroll
                            for (int kk = 0; kk < KZZ; ++kk) {
                                const unsigned zz = (unsigned)(zlz + ZLNZ * kk);
                                if (zz - zoff < zlen) ez[kk] = __ldg(tables + ((int)r.x + (int)zz));
                            }
                        }
                    };
                    double ezA[KZZ], ezB[KZZ];
                    ldez(0, ezA);
                    ldez(1, ezB);
                    // products EX[i] * EY[j] (* C[i][j]) of every (record, column) of the batch, two instructions' worth in flight
                    for (int k0 = prs; k0 < m + prs; k0 += 2 * RPI) {       // (warp-uniform trip count)
                        const int k1 = k0 + RPI;
                        const bool v0 = k0 < m, v1 = k1 < m;
                        double x0 = 0.0, y0 = 0.0, c0 = 1.0, x1 = 0.0, y1 = 0.0, c1 = 1.0;
                        if (v0) {
                            const PairRec r = rbuf[k0];
                            x0 = __ldg(tables + ((int)r.y + pi)); y0 = __ldg(tables + ((int)r.z + pcy));
                            if (MODE == SPLAT_MONO) { const PairAux a = abuf[k0]; c0 = __ldg(tt.ctab + ((int)a.x + pi * (int)a.y + pcy)); }
                        }
                        if (v1) {
                            const PairRec r = rbuf[k1];
                            x1 = __ldg(tables + ((int)r.y + pi)); y1 = __ldg(tables + ((int)r.z + pcy));
                            if (MODE == SPLAT_MONO) { const PairAux a = abuf[k1]; c1 = __ldg(tt.ctab + ((int)a.x + pi * (int)a.y + pcy)); }
                        }
                        if (v0) { double e = x0 * y0; if (MODE == SPLAT_MONO) e *= c0; pbuf[k0 * NCOL + pc] = e; }
                        if (v1) { double e = x1 * y1; if (MODE == SPLAT_MONO) e *= c1; pbuf[k1 * NCOL + pc] = e; }
                    }
                    __syncwarp();
                    auto accumulate = [&](int k, const double (&ez)[KZZ]) {
                        const double2* e2 = reinterpret_cast<const double2*>(pbuf + k * NCOL + cg * NCL);
#pragma unroll
                        for (int c = 0; c < NCL; c += 2) {
                            const double2 e = e2[c >> 1];
#pragma unroll
                            for (int kk = 0; kk < KZZ; ++kk) {
                                acc[c][kk] += __double_as_longlong(__fma_rn(e.x, ez[kk], MDSF_MAGIC)) - MDSF_MAGIC_BITS;
                                acc[c + 1][kk] += __double_as_longlong(__fma_rn(e.y, ez[kk], MDSF_MAGIC)) - MDSF_MAGIC_BITS;
                            }
                        }
                    };
                    for (int k = 0; k < m; k += 2) {
                        accumulate(k, ezA);
                        ldez(k + 2, ezA);
                        if (k + 1 < m) { accumulate(k + 1, ezB); ldez(k + 3, ezB); }
                    }
                }
                b += RBZ;
            } while (b < n);
            // fixed point -> fp64 (overflow: a cell held > 2048 peak amplitudes)
            long long any = 0;
#pragma unroll
            for (int kk = 0; kk < KZZ; ++kk) {
                const int z = s * ZW + zlz + ZLNZ * kk;
                double* cell = smem + tile_index(gp, NCOL, cg * NCL, z, part);
                const int cstride = gp.zilv ? 2 * nzp : nzp;            // doubles between consecutive columns
#pragma unroll
                for (int c = 0; c < NCL; ++c) {
                    any |= acc[c][kk];
                    if (z < nz) cell[(size_t)c * cstride] = fx_to_double(acc[c][kk]) * gp.fx_inv;
                }
            }
            ovf |= (any >> 62) != 0;
            item = item2; lbeg = lbeg2; n = n2;
        }
    } else {
    // work items: (group of SUB consecutive slabs, part); the first MDSF_SPLAT_WARPS go out statically, the rest are
    // claimed from a shared counter (list lengths vary: the barrier before the FFT waited 15% of the time for the longest)
    const int nsg = (gp.nslab + SUB - 1) / SUB, nitems = 2 * nsg;
    for (int item = warp; item < nitems;) {
        const int sg = item >> 1, part = item & 1;
        const int s = sg * SUB + grp;                         // my group's slab
        const bool slab_ok = s < gp.nslab;
        const int zbase = s * ZW + zl;                        // my cells: x = X0 + i, y = Y0 + cy, z = zbase + ZLN * k (lanes walk z: conflict-free stores)
        {
            const int f = 2 * q + part;
            const int cstride = gp.zilv ? 2 * nzp : nzp;                        // doubles between consecutive columns
            if (MODE == SPLAT_DENSITY) {
#pragma unroll
                for (int i = 0; i < TX; ++i) {
                    const bool live = f < nframes && X0 + i < gp.n[0] && Y0 + cy < gp.n[1];
                    const double* src = src_density + (((long long)f * gp.n[0] + (X0 + i)) * gp.n[1] + (Y0 + cy)) * nz;
#pragma unroll
                    for (int k = 0; k < KZ; ++k) {
                        const int z = zbase + ZLN * k;
                        if (slab_ok && z < nz) smem[tile_index(gp, NCOL, i * TY + cy, z, part)] = live ? src[z] : 0.0;
                    }
                }
            } else {
            long long acc[TX][KZ];
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int k = 0; k < KZ; ++k) acc[i][k] = 0;
            const unsigned key = (unsigned)(f * ntiles + tile) * (unsigned)gp.nslab + (unsigned)(slab_ok ? s : 0);
            const unsigned lbeg = key ? start[key - 1] : 0u;                    // K2 leaves the list ENDS in `start`
            const int n = slab_ok ? (int)(start[key] - lbeg) : 0;               // my group's list
            int nmax = n;                                                       // the longest list of the warp
            if (SUB > 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
            PairRec nxt = make_uint4(0u, 0u, 0u, 0u);
            PairAux nxa = make_uint2(0u, 0u);
            if (gl < RB && gl < n) { nxt = prec[2 * (size_t)(lbeg + gl)]; if (MODE != SPLAT_ORTHO) { const uint4 t = prec[2 * (size_t)(lbeg + gl) + 1]; nxa = make_uint2(t.x, t.y); } }
            for (int b = 0; b < nmax; b += RB) {
                const int m = min(RB, n - b);                  // my group's records in this batch (<= 0: none)
                const int mmax = min(RB, nmax - b);
                __syncwarp();                                  // the previous batch is consumed
                if (gl < RB) { rbuf[grp * RB + gl] = nxt; if (MODE != SPLAT_ORTHO) abuf[grp * RB + gl] = nxa; }
                __syncwarp();
                if (gl < RB && b + RB + gl < n) { nxt = prec[2 * (size_t)(lbeg + b + RB + gl)]; if (MODE != SPLAT_ORTHO) { const uint4 t = prec[2 * (size_t)(lbeg + b + RB + gl) + 1]; nxa = make_uint2(t.x, t.y); } }
                if (MODE == SPLAT_GENERAL) {
                    // one exp per cell: amp * exp(-|b . ucell|^2 / (2 sigma^2))  (dens.py:299-308)
                    for (int r_i = 0; r_i < m; ++r_i) {
                        const PairRec r = rbuf[grp * RB + r_i];
                        const unsigned g = r.w;
                        const int cx0 = g & 7, cx1 = (g >> 3) & 15, cy0 = (g >> 7) & 7, cy1 = (g >> 10) & 15;
                        const int zoff = (g >> 14) & 127, zend = (g >> 21) & 127;
                        if (cy < cy0 || cy >= cy1) continue;
                        const AtomRec* ar = recs + (long long)f * gp.natoms + abuf[grp * RB + r_i].x;
                        const double rx = ar->r[0], ry = ar->r[1], rz = ar->r[2];
                        const int type = ar->type;
                        const double by = __dsub_rn(ry, __dmul_rn((double)((int)r.z + cy), gp.dr[1]));
                        const double t2 = tt.two_sig2[type], amp = tt.amp[type];
#pragma unroll
                        for (int i = 0; i < TX; ++i) {
                            if (i < cx0 || i >= cx1) continue;
                            const double bx = __dsub_rn(rx, __dmul_rn((double)((int)r.y + i), gp.dr[0]));
#pragma unroll
                            for (int k = 0; k < KZ; ++k) {
                                const int zz = zl + ZLN * k;
                                if (zz < zoff || zz >= zend) continue;
                                const double bzv = __dsub_rn(rz, __dmul_rn((double)((int)r.x + zz), gp.dr[2]));
                                const double c0 = gp.u[0] * bx + gp.u[3] * by + gp.u[6] * bzv;
                                const double c1 = gp.u[1] * bx + gp.u[4] * by + gp.u[7] * bzv;
                                const double c2 = gp.u[2] * bx + gp.u[5] * by + gp.u[8] * bzv;
                                acc[i][k] += __double2ll_rn(amp * exp(-(c0 * c0 + c1 * c1 + c2 * c2) / t2) * gp.fx_scale);
                            }
                        }
                    }
                    continue;
                }
                // ---- separable ucell: cp.async staging ring, NS-1 records ahead.  A group whose list is exhausted
                // stages nothing and accumulates the zeros of a slot it cleared once (below).
                auto produce = [&](int i, int slot) {
                    if (i >= m) return;
                    const PairRec r = rbuf[grp * RB + i];
                    const unsigned dst = slots_s + (unsigned)((slot * SUB + grp) * SLOTG + gl) * 8u;
#pragma unroll
                    for (int rd = 0; rd < ROUNDS; ++rd) {
                        if ((rd + 1) * GL <= ZW) {             // every lane stages an EZ entry
                            const unsigned e = (unsigned)(gl + GL * rd);
                            cp_async8_zfill(dst + rd * GL * 8u, tables + ((int)r.x + (int)e), e - (r.w & 127u) < (r.w >> 7));
                        } else {
                            const int e = gl + GL * rd;
                            if (e < ZW) {
                                cp_async8_zfill(dst + rd * GL * 8u, tables + ((int)r.x + e), (unsigned)e - (r.w & 127u) < (r.w >> 7));
                            } else if (skind[rd] != 0) {
                                const double* src;
                                if (MODE == SPLAT_MONO && skind[rd] == 3) {
                                    const PairAux a = abuf[grp * RB + i];
                                    src = tt.ctab + ((int)a.x + smul[rd] * (int)a.y + soff[rd]);
                                } else {
                                    src = tables + ((int)(skind[rd] == 1 ? r.y : r.z) + soff[rd]);
                                }
                                cp_async8(dst + rd * GL * 8u, src);
                            }
                        }
                    }
                };
                int ps = 0, cs = 0;
#pragma unroll
                for (int i = 0; i < NS - 1; ++i) {
                    produce(i, ps);
                    cp_async_commit();
                    ps = (ps + 1 == NS) ? 0 : ps + 1;
                }
                for (int r_i = 0; r_i < mmax; ++r_i) {
                    produce(r_i + NS - 1, ps);
                    cp_async_commit();
                    ps = (ps + 1 == NS) ? 0 : ps + 1;
                    cp_async_wait<NS - 1>();
                    __syncwarp();
                    if (SUB == 1 || r_i < m) {
                    const double* sl = slots + (cs * SUB + grp) * SLOTG;
                    const double ey = sl[ZW + TX + cy];
                    double exy[TX], ez[KZ];
                    const double2* ex2 = reinterpret_cast<const double2*>(sl + ZW);                        // EX: the same address in every lane of the group
                    const double2* cc2 = reinterpret_cast<const double2*>(sl + ZW + TX + TY + cy * TX);    // my row of the cross-term table
#pragma unroll
                    for (int i = 0; i < TX; i += 2) {
                        const double2 e = ex2[i >> 1];
                        exy[i] = e.x * ey; exy[i + 1] = e.y * ey;
                        if (MODE == SPLAT_MONO) { const double2 cv = cc2[i >> 1]; exy[i] *= cv.x; exy[i + 1] *= cv.y; }
                    }
#pragma unroll
                    for (int k = 0; k < KZ; ++k) ez[k] = sl[zl + ZLN * k];
#pragma unroll
                    for (int i = 0; i < TX; ++i)
#pragma unroll
                        for (int k = 0; k < KZ; ++k)
                            acc[i][k] += __double_as_longlong(__fma_rn(exy[i], ez[k], MDSF_MAGIC)) - MDSF_MAGIC_BITS;
                    }
                    cs = (cs + 1 == NS) ? 0 : cs + 1;
                    __syncwarp();                              // slot cs-1 may be refilled
                }
            }
            // fixed point -> fp64 (overflow: a cell held > 2048 peak amplitudes)
            long long any = 0;
#pragma unroll
            for (int k = 0; k < KZ; ++k) {
                const int z = zbase + ZLN * k;
                const bool live = slab_ok && z < nz;
                double* cell = smem + tile_index(gp, NCOL, cy, z, part);             // column (i, cy) = i * TY + cy
#pragma unroll
                for (int i = 0; i < TX; ++i) {
                    any |= acc[i][k];
                    if (live) cell[i * TY * cstride] = fx_to_double(acc[i][k]) * gp.fx_inv;
                }
            }
            ovf |= (any >> 62) != 0;                           // terms are >= 0: bit 62 or 63 set in any sum
            }
        }
        int nxt_item = 0;
        if (lane == 0) nxt_item = atomicAdd(&s_next, 1);
        item = __shfl_sync(0xffffffffu, nxt_item, 0);
    }
    }
    if (ovf) atomicExch(err_flag, 2);
    cp_async_wait<0>();
    __syncthreads();

    double* twr = reinterpret_cast<double*>(area);            // generic path / late load: twiddles reuse the staging area
    double* twi = twr + nz;
    if (zspec && tws_off == 0) {
        tws_s = reinterpret_cast<double2*>(area);
        for (int i = threadIdx.x; i < tws_n; i += MDSF_SPLAT_THREADS) tws_s[i] = tws_g[i];
    } else if (FUSE && !zspec) {
        load_twiddles(twr, twi, twz, nz);
    }
    if (dens_dump != nullptr) {                               // parity tap: plain [q][x][y][z] pairs, before any FFT
        for (int i = threadIdx.x; i < NCOL * nz; i += blockDim.x) {
            const int cc = i / nz, z = i - cc * nz;
            const int x = X0 + (cc >> LTY), y = Y0 + (cc & (TY - 1));
            if (x < gp.n[0] && y < gp.n[1]) {
                dens_dump[(((long long)q * gp.n[0] + x) * gp.n[1] + y) * nz + z] =
                    make_double2(smem[tile_index(gp, NCOL, cc, z, 0)], smem[tile_index(gp, NCOL, cc, z, 1)]);
            }
        }
    }
    __syncthreads();

    const long long ncell = (long long)gp.n[0] * gp.n[1] * nz;
    const long long cs_ = (long long)gp.n[0] * gp.n[1] * gp.lw;      // chunk stride
    double2* volq = vol + (long long)q * ncell;
    if (FUSE) {
        // the tile size a grid selects fixes its z length class: compile-time stages for the power-of-two lengths
        // (and 768) the BASELINE configs use, the generic run-time stages for everything else
        if (!zspec) fft_tile_z(tile_re, tile_im, twr, twi, zplan, NCOL, nzp, gp.pad_shift);      // ends with a barrier
        else zfft2<NCOL>(reinterpret_cast<double2*>(smem), tws_s, nzp, nz);
    }
    // store: every tile row x is one run of TY * Nz contiguous cells (plain layout), or Nz / lw runs of TY * lw cells
    // (chunked layout): 16-byte stores, consecutive threads -> consecutive addresses.  One flat index over the tile,
    // 32-bit arithmetic inside the tile row.
    const bool pw2 = (nz & (nz - 1)) == 0;
    const int lnz = __ffs(nz) - 1, llw = __ffs(gp.lw) - 1;
    const int per_x = TY * nz, xvalid = min(TX, gp.n[0] - X0), yvalid = min(TY, gp.n[1] - Y0);
    const bool chunked = gp.lw != nz;
    const int rowlen = gp.n[1] * gp.lw;
    double2* dst0 = volq + (long long)X0 * rowlen + (long long)Y0 * gp.lw;
    for (int i = threadIdx.x; i < xvalid * per_x; i += MDSF_SPLAT_THREADS) {
        int xx, j;
        if (pw2) { xx = i >> (lnz + LTY); j = i & (per_x - 1); } else { xx = i / per_x; j = i - xx * per_x; }
        int yy, z;
        long long o;
        if (!chunked) {                                        // plain: j = yy * Nz + z
            if (pw2) { yy = j >> lnz; z = j & (nz - 1); } else { yy = j / nz; z = j - yy * nz; }
            o = (long long)xx * rowlen + j;
        } else {                                               // chunked: j = (ch * TY + yy) * lw + zw
            const int ch = j >> (LTY + llw), zw = j & (gp.lw - 1);
            yy = (j >> llw) & (TY - 1);
            z = (ch << llw) + zw;
            o = (long long)ch * cs_ + (long long)(xx * rowlen + yy * gp.lw + zw);
        }
        if (yy >= yvalid) continue;
        if (gp.zilv) {
            dst0[o] = reinterpret_cast<const double2*>(smem)[((xx << LTY) + yy) * nzp + z];
        } else {
            const int a = ((xx << LTY) + yy) * nzp + zpos(z, gp.pad_shift);
            dst0[o] = make_double2(tile_re[a], tile_im[a]);
        }
    }
}
