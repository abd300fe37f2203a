// Instantiations + launcher of splat_zfft_kernel (see mdsf_splat.cuh).
#include "mdsf_launch.h"
#include "mdsf_splat.cuh"

static const int kMaxSmemSplat = 227 * 1024 - 1536;     // the kernels also hold up to ~1.1 KB of static shared memory (item counter, list bounds)

template <int LCOL> static size_t warp_bytes(int sub) { return sub == 2 ? (size_t)SplatGeom<LCOL, 2>::WARP_BYTES : (size_t)SplatGeom<LCOL, 1>::WARP_BYTES; }

size_t mdsf_splat_smem(int lcol, int sub, int nzp, int nz, int zlane) {
    size_t wb = lcol == 2 ? warp_bytes<2>(sub) : (lcol == 3 ? warp_bytes<3>(sub) : (lcol == 4 ? warp_bytes<4>(sub) : warp_bytes<5>(sub)));
    if (zlane) {       // (the density mode of a z-lane handle still uses the SUB = 1 staging layout: size for both)
        const size_t zb = lcol == 2 ? ZLaneGeom<2>::WARP_BYTES : (lcol == 3 ? ZLaneGeom<3>::WARP_BYTES : (lcol == 4 ? ZLaneGeom<4>::WARP_BYTES : ZLaneGeom<5>::WARP_BYTES));
        if (zb > wb) wb = zb;
    }
    size_t area = wb * MDSF_SPLAT_WARPS;
    const size_t tw = (size_t)2 * (nz > 256 ? nz : 256) * sizeof(double);     // z twiddles reuse the staging area
    if (area < tw) area = tw;
    return (size_t)2 * ((size_t)1 << lcol) * nzp * sizeof(double) + area;
}

bool mdsf_zspec_length(int nz) { return zspec_length(nz); }

template <int LCOL, int MODE>
static cudaError_t launch1(bool fuse, dim3 grid, size_t smem, cudaStream_t st, const SplatArgs& a) {
    if (a.gp.zlane && (MODE == SPLAT_ORTHO || MODE == SPLAT_MONO))
        splat_zfft_kernel<LCOL, MODE, 1, true><<<grid, MDSF_SPLAT_THREADS, smem, st>>>(a.prec, a.start, a.recs, a.tables, a.src_density, a.nframes,
            a.vol, a.dens_dump, a.gp, a.tt, a.zplan, a.twz, a.err_flag, fuse ? 1 : 0, a.tws, a.tws_n, a.tws_off);
    else if (a.gp.sub == 2)
        splat_zfft_kernel<LCOL, MODE, 2><<<grid, MDSF_SPLAT_THREADS, smem, st>>>(a.prec, a.start, a.recs, a.tables, a.src_density, a.nframes,
            a.vol, a.dens_dump, a.gp, a.tt, a.zplan, a.twz, a.err_flag, fuse ? 1 : 0, a.tws, a.tws_n, a.tws_off);
    else
        splat_zfft_kernel<LCOL, MODE, 1><<<grid, MDSF_SPLAT_THREADS, smem, st>>>(a.prec, a.start, a.recs, a.tables, a.src_density, a.nframes,
            a.vol, a.dens_dump, a.gp, a.tt, a.zplan, a.twz, a.err_flag, fuse ? 1 : 0, a.tws, a.tws_n, a.tws_off);
    return cudaGetLastError();
}

template <int LCOL>
static cudaError_t launch_mode(int mode, bool fuse, dim3 grid, size_t smem, cudaStream_t st, const SplatArgs& a) {
    switch (mode) {
        case SPLAT_ORTHO:   return launch1<LCOL, SPLAT_ORTHO>(fuse, grid, smem, st, a);
        case SPLAT_MONO:    return launch1<LCOL, SPLAT_MONO>(fuse, grid, smem, st, a);
        case SPLAT_GENERAL: return launch1<LCOL, SPLAT_GENERAL>(fuse, grid, smem, st, a);
        case SPLAT_DENSITY: return launch1<LCOL, SPLAT_DENSITY>(fuse, grid, smem, st, a);
    }
    return cudaErrorInvalidValue;
}

cudaError_t mdsf_launch_splat(int lcol, int mode, bool fuse, dim3 grid, size_t smem, cudaStream_t st, const SplatArgs& a) {
    switch (lcol) {
        case 2: return launch_mode<2>(mode, fuse, grid, smem, st, a);
        case 3: return launch_mode<3>(mode, fuse, grid, smem, st, a);
        case 4: return launch_mode<4>(mode, fuse, grid, smem, st, a);
        case 5: return launch_mode<5>(mode, fuse, grid, smem, st, a);
    }
    return cudaErrorInvalidValue;
}

template <int LCOL, int MODE> static cudaError_t cfg1() {
    cudaError_t e = cudaFuncSetAttribute(splat_zfft_kernel<LCOL, MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemSplat);
    if (e != cudaSuccess) return e;
    if (MODE == SPLAT_ORTHO || MODE == SPLAT_MONO) {
        e = cudaFuncSetAttribute(splat_zfft_kernel<LCOL, MODE, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemSplat);
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(splat_zfft_kernel<LCOL, MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemSplat);
}
template <int LCOL> static cudaError_t cfg_mode() {
    cudaError_t e;
    if ((e = cfg1<LCOL, SPLAT_ORTHO>()) != cudaSuccess) return e;
    if ((e = cfg1<LCOL, SPLAT_MONO>()) != cudaSuccess) return e;
    if ((e = cfg1<LCOL, SPLAT_GENERAL>()) != cudaSuccess) return e;
    return cfg1<LCOL, SPLAT_DENSITY>();
}
cudaError_t mdsf_splat_configure(void) {
    cudaError_t e;
    if ((e = cfg_mode<2>()) != cudaSuccess) return e;
    if ((e = cfg_mode<3>()) != cudaSuccess) return e;
    if ((e = cfg_mode<4>()) != cudaSuccess) return e;
    return cfg_mode<5>();
}
