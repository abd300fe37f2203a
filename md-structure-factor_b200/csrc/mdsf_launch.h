// Host-callable launchers of the kernel translation units (mdsf_splat_inst.cu, mdsf_pass_inst.cu), so the
// template instantiations compile in parallel with the C ABI file.
#pragma once
#include <cuda_runtime.h>
#include "mdsf_common.cuh"
#include "mdsf_fft.cuh"

struct SplatArgs {
    const uint4* prec /* 32-byte records */; const unsigned* start; const AtomRec* recs; const double* tables;
    const double* src_density; int nframes;
    double2* vol; double2* dens_dump; GridParams gp; TypeTable tt; FftPlan zplan; const double2* twz; int* err_flag;
    const double2* tws; int tws_n; int tws_off;      // per-stage z twiddle tables (compile-time z path): device pointer, entries, smem byte offset (0: late load)
};
// mode: SPLAT_ORTHO / SPLAT_MONO / SPLAT_GENERAL / SPLAT_DENSITY (mdsf_splat.cuh); grid = (tiles, pairs)
cudaError_t mdsf_launch_splat(int lcol, int mode, bool fuse, dim3 grid, size_t smem, cudaStream_t st, const SplatArgs& a);
cudaError_t mdsf_splat_configure(void);                   // opt in to the large dynamic shared memory sizes
size_t mdsf_splat_smem(int lcol, int sub, int nzp, int nz, int zlane = 0);        // dynamic shared memory of one splat CTA (without a twiddle region)
bool mdsf_zspec_length(int nz);                           // compile-time z stages (interleaved tile) exist for this z length

struct PassArgs {
    double2* vol; double* P; const FftPlan* plan; const double2* tw; PassGeom pg; int npairs; int nouter;   // nouter: Nx (y pass) / Ny (x pass)
    int ntile;                  // z tiles (grid.x)
    double2* scratch;           // fused y->x path: L2-resident hand-over buffer (or nullptr)
};
cudaError_t mdsf_pass_configure(void);
// returns the number of launches issued (>= 1) or -1 with the CUDA error in *err
int mdsf_launch_pass_y(const PassArgs& a, cudaStream_t st, cudaError_t* err);
int mdsf_launch_pass_x(const PassArgs& a, cudaStream_t st, cudaError_t* err);

// fused y -> x pass (mdsf_yx.cuh): one persistent kernel per batch, hand-over through L2
struct YXParams;
bool mdsf_yx_supported(int ny, int nx);
int mdsf_yx_blocks_per_sm(int ny, int nx);
cudaError_t mdsf_launch_yx(int ny, int nx, const YXParams& p, int grid, cudaStream_t st);

// TMA-fed persistent y / x passes of 512-point axes on the lw = 8 chunked layout (mdsf_tma_pass.cuh)
struct TPParams;
cudaError_t mdsf_tma_pass_configure(void);
cudaError_t mdsf_launch_tma_pass(bool xpass, const TPParams& p, int grid, cudaStream_t st);
