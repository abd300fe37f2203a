// Instantiations + launcher of splat_zfft_kernel (see mdsf_splat.cuh).
#include "mdsf_launch.h"
#include "mdsf_splat.cuh"

static const int kMaxSmemSplat = 227 * 1024;

template <int LCOL> static size_t warp_bytes() { return (size_t)SplatGeom<LCOL>::WARP_BYTES; }

size_t mdsf_splat_smem(int lcol, int nzp, int nz) {
    size_t wb = lcol == 2 ? warp_bytes<2>() : (lcol == 3 ? warp_bytes<3>() : (lcol == 4 ? warp_bytes<4>() : warp_bytes<5>()));
    size_t area = wb * MDSF_SPLAT_WARPS;
    const size_t tw = (size_t)2 * (nz > 256 ? nz : 256) * sizeof(double);     // z twiddles reuse the staging area
    if (area < tw) area = tw;
    return (size_t)2 * ((size_t)1 << lcol) * nzp * sizeof(double) + area;
}

template <int LCOL, int MODE>
static cudaError_t launch1(bool fuse, dim3 grid, size_t smem, cudaStream_t st, const SplatArgs& a) {
    splat_zfft_kernel<LCOL, MODE><<<grid, MDSF_SPLAT_THREADS, smem, st>>>(a.prec, a.paux, a.start, a.recs, a.tables, a.src_density, a.nframes,
        a.vol, a.dens_dump, a.gp, a.tt, a.zplan, a.twz, a.err_flag, fuse ? 1 : 0);
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
    return cudaFuncSetAttribute(splat_zfft_kernel<LCOL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemSplat);
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
