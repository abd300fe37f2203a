// Instantiations + dispatch of the y / x FFT passes (see mdsf_fft.cuh).
#include "mdsf_launch.h"
#include "mdsf_yx.cuh"
#include "mdsf_tma_pass.cuh"

static const int kMaxSmemPass = 227 * 1024;

#ifndef MDSF_FFT3_MINB_Y
#define MDSF_FFT3_MINB_Y 2     // CTAs per SM of the three-stage y pass (measured: 2 -> 4.87 ms, 3 -> 4.95 ms on c3)
#endif
#ifndef MDSF_FFT3_MINB_X
#define MDSF_FFT3_MINB_X 2     // ... of the x pass (register accumulators: 116 registers)
#endif

#define SETATTR(k)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPass); \
        if (e_ != cudaSuccess) return e_;                                                            \
    } while (0)

cudaError_t mdsf_pass_configure(void) {
    SETATTR((fft_y_kernel<128, MDSF_PASS_MINBLOCKS>));
    SETATTR((fft_y_kernel<256, 2>));
    SETATTR((fft_y_kernel<512, 1>));
    SETATTR((fft_x_accum_kernel<128, MDSF_PASS_MINBLOCKS>));
    SETATTR((fft_x_accum_kernel<256, 2>));
    SETATTR((fft_x_accum_kernel<512, 1>));
    SETATTR((fft_y_fast_kernel<16, 16>));
    SETATTR((fft_y_fast_kernel<8, 8>));
    SETATTR((fft_x_accum_async_kernel<16, 16>));
    SETATTR((fft_x_accum_async_kernel<8, 8>));
    SETATTR((fft3_pass_kernel<16, 16, 3, 3, 256, 2, false>));
    SETATTR((fft3_pass_kernel<16, 16, 3, 3, 256, 2, true>));
    SETATTR((fft3_pass_kernel<8, 8, 16, 3, 256, 1, false>));
    SETATTR((fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_Y, false>));
    SETATTR((fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_X, true>));
    return cudaSuccess;
}

static int ilog2(int w) { int l = 0; while ((1 << l) < w) ++l; return l; }

// threads of the generic kernels: [n][W] tile(s) + twiddles decide how many CTAs fit an SM
static int generic_threads(int n, int W, int arrays) {
    const size_t one = (size_t)n * W * 8 * arrays + (size_t)n * 16;
    if (one <= 100 * 1024) return 128;
    return one <= 110 * 1024 ? 256 : 512;
}

int mdsf_launch_pass_y(const PassArgs& a, cudaStream_t st, cudaError_t* err) {
    const FftPlan& p = *a.plan;
    const int n = p.n, W = a.pg.W, logw = ilog2(W);
    const size_t sm = (size_t)2 * n * W * 8 + (size_t)2 * n * 8;
    dim3 grid(a.ntile, a.nouter, a.npairs);
    const bool two = p.nstages == 2 && p.radix[0] == p.radix[1] && (n / p.radix[0]) * W == MDSF_PASS_THREADS;
    const bool three = p.nstages == 3 && W == 8;
    if (two && p.radix[0] == 16)
        fft_y_fast_kernel<16, 16><<<grid, MDSF_PASS_THREADS, sm, st>>>(a.vol, a.tw, a.pg, logw);
    else if (two && p.radix[0] == 8)
        fft_y_fast_kernel<8, 8><<<grid, MDSF_PASS_THREADS, sm, st>>>(a.vol, a.tw, a.pg, logw);
    else if (three && p.radix[0] == 8 && p.radix[1] == 8 && p.radix[2] == 8)
        fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_Y, false><<<grid, 256, sm, st>>>(a.vol, nullptr, a.tw, a.pg, a.npairs);
    else if (three && p.radix[0] == 16 && p.radix[1] == 16 && p.radix[2] == 3)
        fft3_pass_kernel<16, 16, 3, 3, 256, 2, false><<<grid, 256, sm, st>>>(a.vol, nullptr, a.tw, a.pg, a.npairs);
    else if (three && p.radix[0] == 8 && p.radix[1] == 8 && p.radix[2] == 16)
        fft3_pass_kernel<8, 8, 16, 3, 256, 1, false><<<grid, 256, sm, st>>>(a.vol, nullptr, a.tw, a.pg, a.npairs);
    else {
        const int thr = generic_threads(n, W, 2);
        if (thr == 512) fft_y_kernel<512, 1><<<grid, 512, sm, st>>>(a.vol, p, a.tw, a.pg, logw);
        else if (thr == 256) fft_y_kernel<256, 2><<<grid, 256, sm, st>>>(a.vol, p, a.tw, a.pg, logw);
        else fft_y_kernel<128, MDSF_PASS_MINBLOCKS><<<grid, 128, sm, st>>>(a.vol, p, a.tw, a.pg, logw);
    }
    *err = cudaGetLastError();
    return *err == cudaSuccess ? 1 : -1;
}

int mdsf_launch_pass_x(const PassArgs& a, cudaStream_t st, cudaError_t* err) {
    const FftPlan& p = *a.plan;
    const int n = p.n, W = a.pg.W, logw = ilog2(W);
    const size_t sm3 = (size_t)3 * n * W * 8 + (size_t)2 * n * 8;      // generic kernel: + |C|^2 tile
    const size_t sm2 = (size_t)2 * n * W * 8 + (size_t)2 * n * 8;
    dim3 grid(a.ntile, a.nouter);
    const bool two = p.nstages == 2 && p.radix[0] == p.radix[1] && (n / p.radix[0]) * W == MDSF_PASS_THREADS;
    const bool three = p.nstages == 3 && W == 8;
    const size_t sma = sm2 + (size_t)p.radix[0] * MDSF_PASS_THREADS * 16;     // + cp.async staging slots
    if (two && p.radix[0] == 16)
        fft_x_accum_async_kernel<16, 16><<<grid, MDSF_PASS_THREADS, sma, st>>>(a.vol, a.P, a.tw, a.pg, logw, a.npairs);
    else if (two && p.radix[0] == 8)
        fft_x_accum_async_kernel<8, 8><<<grid, MDSF_PASS_THREADS, sma, st>>>(a.vol, a.P, a.tw, a.pg, logw, a.npairs);
    else if (three && p.radix[0] == 8 && p.radix[1] == 8 && p.radix[2] == 8)
        fft3_pass_kernel<8, 8, 8, 3, 256, MDSF_FFT3_MINB_X, true><<<grid, 256, sm2, st>>>(a.vol, a.P, a.tw, a.pg, a.npairs);
    else if (three && p.radix[0] == 16 && p.radix[1] == 16 && p.radix[2] == 3)
        fft3_pass_kernel<16, 16, 3, 3, 256, 2, true><<<grid, 256, sm2, st>>>(a.vol, a.P, a.tw, a.pg, a.npairs);
    else {
        // (1024-point x axes stay on the generic kernel: the three-stage variant measured slower, 19.8 -> 25.2 ms on c5)
        const int thr = generic_threads(n, W, 3);
        if (thr == 512) fft_x_accum_kernel<512, 1><<<grid, 512, sm3, st>>>(a.vol, a.P, p, a.tw, a.pg, logw, a.npairs);
        else if (thr == 256) fft_x_accum_kernel<256, 2><<<grid, 256, sm3, st>>>(a.vol, a.P, p, a.tw, a.pg, logw, a.npairs);
        else fft_x_accum_kernel<128, MDSF_PASS_MINBLOCKS><<<grid, 128, sm3, st>>>(a.vol, a.P, p, a.tw, a.pg, logw, a.npairs);
    }
    *err = cudaGetLastError();
    return *err == cudaSuccess ? 1 : -1;
}

// ---- fused y -> x pass (mdsf_yx.cuh)
static const int kYxSmem = 2 * (512 + 64) * MDSF_YX_W * 16;      // two padded [512][4] tile buffers

template <int NY, int NX> static cudaError_t yx_go(const YXParams& p, int grid, cudaStream_t st) {
    yx_pass_kernel<NY, NX><<<grid, MDSF_YX_CTA, kYxSmem, st>>>(p);
    return cudaGetLastError();
}
template <int NY, int NX> static int yx_occ() {
    int per_sm = 0;
    if (cudaFuncSetAttribute(yx_pass_kernel<NY, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, kYxSmem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, yx_pass_kernel<NY, NX>, MDSF_YX_CTA, kYxSmem) != cudaSuccess) return 0;
    return per_sm;
}
bool mdsf_yx_supported(int ny, int nx) { return (ny == 256 || ny == 512) && (nx == 256 || nx == 512); }
int mdsf_yx_blocks_per_sm(int ny, int nx) {
    if (ny == 512 && nx == 512) return yx_occ<512, 512>();
    if (ny == 256 && nx == 256) return yx_occ<256, 256>();
    if (ny == 512 && nx == 256) return yx_occ<512, 256>();
    if (ny == 256 && nx == 512) return yx_occ<256, 512>();
    return 0;
}
cudaError_t mdsf_launch_yx(int ny, int nx, const YXParams& p, int grid, cudaStream_t st) {
    if (ny == 512 && nx == 512) return yx_go<512, 512>(p, grid, st);
    if (ny == 256 && nx == 256) return yx_go<256, 256>(p, grid, st);
    if (ny == 512 && nx == 256) return yx_go<512, 256>(p, grid, st);
    if (ny == 256 && nx == 512) return yx_go<256, 512>(p, grid, st);
    return cudaErrorInvalidValue;
}

// ---- TMA-fed persistent passes for 512-point axes (mdsf_tma_pass.cuh)
cudaError_t mdsf_tma_pass_configure(void) {
    cudaError_t e = cudaFuncSetAttribute(tma_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MDSF_TP_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tma_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MDSF_TP_SMEM);
}
cudaError_t mdsf_launch_tma_pass(bool xpass, const TPParams& p, int grid, cudaStream_t st) {
    if (xpass) tma_pass_kernel<true><<<grid, MDSF_TP_CTA, MDSF_TP_SMEM, st>>>(p);
    else tma_pass_kernel<false><<<grid, MDSF_TP_CTA, MDSF_TP_SMEM, st>>>(p);
    return cudaGetLastError();
}
