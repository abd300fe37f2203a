// Hand-written fp64 FFT passes for sm_100a (replaces np.fft.rfftn, dens.py:313).
//
// Two real frames travel as one complex volume (re = frame 2q, im = frame 2q+1), so every pass
// is a plain complex transform; |A|^2+|B|^2 of the two real transforms is recovered from
// (|C(k)|^2 + |C(-k)|^2)/2 when S(q) is read out.  Each pass is an in-place decimation-in-
// frequency FFT over one axis of a shared-memory tile: a stage loads R points into registers,
// applies an R-point DFT and the inter-stage twiddles, and writes the R results back to the
// SAME addresses, so stages only need a barrier between them.  The output of a DIF transform
// is digit-reversed; nothing is reordered per frame -- the volume and the |C|^2 accumulator
// stay in "position space" on all three axes and mdsf_read_sf() applies the three digit
// reversal tables once.  There is no tensor-core work here: nothing on this path is a dense
// contraction, the passes are bound by shared-memory/HBM traffic and the fp64 pipe.
#pragma once
#include "mdsf_common.cuh"

struct FftPlan {
    int n;
    int nstages;
    int radix[MDSF_MAX_RADIX_STAGES];
};

#define MDSF_SQRT1_2 0.70710678118654752440
#define MDSF_SQRT3_2 0.86602540378443864676

// ------------------------------------------------------------------ in-register DFTs (forward)
template <int R> struct Dft;

template <> struct Dft<2> {
    __device__ __forceinline__ static void run(double (&xr)[2], double (&xi)[2], const double*, const double*, int) {
        double ar = xr[0], ai = xi[0];
        xr[0] = ar + xr[1]; xi[0] = ai + xi[1];
        xr[1] = ar - xr[1]; xi[1] = ai - xi[1];
    }
};

__device__ __forceinline__ void dft4(double& r0, double& i0, double& r1, double& i1,
                                     double& r2, double& i2, double& r3, double& i3) {
    double ar = r0 + r2, ai = i0 + i2, br = r0 - r2, bi = i0 - i2;
    double cr = r1 + r3, ci = i1 + i3, dr_ = r1 - r3, di = i1 - i3;
    r0 = ar + cr; i0 = ai + ci;
    r2 = ar - cr; i2 = ai - ci;
    r1 = br + di; i1 = bi - dr_;     // b - i d
    r3 = br - di; i3 = bi + dr_;     // b + i d
}

template <> struct Dft<4> {
    __device__ __forceinline__ static void run(double (&xr)[4], double (&xi)[4], const double*, const double*, int) {
        dft4(xr[0], xi[0], xr[1], xi[1], xr[2], xi[2], xr[3], xi[3]);
    }
};

__device__ __forceinline__ void dft8(double* xr, double* xi) {   // xr[0..7] natural in, natural out
    // even / odd 4-point transforms
    dft4(xr[0], xi[0], xr[2], xi[2], xr[4], xi[4], xr[6], xi[6]);   // E0..E3 in slots 0,2,4,6
    dft4(xr[1], xi[1], xr[3], xi[3], xr[5], xi[5], xr[7], xi[7]);   // O0..O3 in slots 1,3,5,7
    double er[4] = {xr[0], xr[2], xr[4], xr[6]}, ei[4] = {xi[0], xi[2], xi[4], xi[6]};
    double or_[4], oi[4];
    or_[0] = xr[1]; oi[0] = xi[1];
    or_[1] = (xr[3] + xi[3]) * MDSF_SQRT1_2; oi[1] = (xi[3] - xr[3]) * MDSF_SQRT1_2;   // * (1-i)/sqrt2
    or_[2] = xi[5]; oi[2] = -xr[5];                                                    // * (-i)
    or_[3] = (xi[7] - xr[7]) * MDSF_SQRT1_2; oi[3] = -(xr[7] + xi[7]) * MDSF_SQRT1_2;  // * (-1-i)/sqrt2
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        xr[k] = er[k] + or_[k]; xi[k] = ei[k] + oi[k];
        xr[k + 4] = er[k] - or_[k]; xi[k + 4] = ei[k] - oi[k];
    }
}

template <> struct Dft<8> {
    __device__ __forceinline__ static void run(double (&xr)[8], double (&xi)[8], const double*, const double*, int) {
        dft8(xr, xi);
    }
};

template <> struct Dft<16> {
    __device__ __forceinline__ static void run(double (&xr)[16], double (&xi)[16], const double*, const double*, int) {
        double er[8], ei[8], or_[8], oi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { er[k] = xr[2 * k]; ei[k] = xi[2 * k]; or_[k] = xr[2 * k + 1]; oi[k] = xi[2 * k + 1]; }
        dft8(er, ei);
        dft8(or_, oi);
        // w16^k = cos(pi k/8) - i sin(pi k/8)
        const double c[8] = {1.0, 0.92387953251128675613, MDSF_SQRT1_2, 0.38268343236508977173,
                             0.0, -0.38268343236508977173, -MDSF_SQRT1_2, -0.92387953251128675613};
        const double s[8] = {0.0, 0.38268343236508977173, MDSF_SQRT1_2, 0.92387953251128675613,
                             1.0, 0.92387953251128675613, MDSF_SQRT1_2, 0.38268343236508977173};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double tr, ti;
            if (k == 0) { tr = or_[0]; ti = oi[0]; }
            else if (k == 4) { tr = oi[4]; ti = -or_[4]; }
            else { tr = or_[k] * c[k] + oi[k] * s[k]; ti = oi[k] * c[k] - or_[k] * s[k]; }
            xr[k] = er[k] + tr; xi[k] = ei[k] + ti;
            xr[k + 8] = er[k] - tr; xi[k + 8] = ei[k] - ti;
        }
    }
};

template <> struct Dft<3> {
    __device__ __forceinline__ static void run(double (&xr)[3], double (&xi)[3], const double*, const double*, int) {
        double tr = xr[1] + xr[2], ti = xi[1] + xi[2];
        double sr = (xr[1] - xr[2]) * MDSF_SQRT3_2, si = (xi[1] - xi[2]) * MDSF_SQRT3_2;
        double mr = xr[0] - 0.5 * tr, mi = xi[0] - 0.5 * ti;
        xr[0] += tr; xi[0] += ti;
        xr[1] = mr + si; xi[1] = mi - sr;    // m - i s
        xr[2] = mr - si; xi[2] = mi + sr;    // m + i s
    }
};

// Odd prime radices (5, 7, 11, 13): direct symmetric DFT; roots come from the axis twiddle
// table (tw[m * n/R] = exp(-2 pi i m / R)).
template <int R> struct Dft {
    __device__ __forceinline__ static void run(double (&xr)[R], double (&xi)[R], const double* twr, const double* twi, int n) {
        constexpr int H = (R - 1) / 2;
        const int step = n / R;
        double sr[H], si[H], dr_[H], di[H];
        double y0r = xr[0], y0i = xi[0];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            sr[j] = xr[j + 1] + xr[R - 1 - j]; si[j] = xi[j + 1] + xi[R - 1 - j];
            dr_[j] = xr[j + 1] - xr[R - 1 - j]; di[j] = xi[j + 1] - xi[R - 1 - j];
            y0r += sr[j]; y0i += si[j];
        }
        double outr[R], outi[R];
        outr[0] = y0r; outi[0] = y0i;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            double ar = xr[0], ai = xi[0], br = 0.0, bi = 0.0;
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const int m = (j * k) % R;
                const double c = twr[m * step], s = -twi[m * step];   // cos, sin of 2 pi m / R
                ar += sr[j - 1] * c; ai += si[j - 1] * c;
                br += dr_[j - 1] * s; bi += di[j - 1] * s;
            }
            outr[k] = ar + bi; outi[k] = ai - br;          // A - i B
            outr[R - k] = ar - bi; outi[R - k] = ai + br;  // A + i B
        }
#pragma unroll
        for (int k = 0; k < R; ++k) { xr[k] = outr[k]; xi[k] = outi[k]; }
    }
};

// ------------------------------------------------------------------ one DIF stage over a tile
// Tile addressing in shared memory: point p of transform f lives at f*fs + pos(p)*es with
// pos(p) = p + (p >> pad).  ZPASS: transforms are contiguous rows (es = 1) and consecutive lanes
// take consecutive butterflies of one row.  Otherwise the tile is [n][W] (fs = 1, es = W, W a
// power of two) and consecutive lanes take consecutive columns, i.e. contiguous global memory.
// A stage may take its input straight from global memory (first stage of a y/x pass) and may
// send its output straight back (last stage), so a 2-stage transform crosses shared memory once.
enum { IO_SMEM = 0, IO_GLOBAL = 1, IO_POWER = 2 };

struct GlobalTile {
    double2* base;        // element (p, f) at base[p * row_stride + f]
    long long row_stride;
    int w;                // valid columns (f < w)
};

template <int R, bool ZPASS, int IN, int OUT>
__device__ __forceinline__ void fft_stage(double* __restrict__ sre, double* __restrict__ sim,
                                          const double* __restrict__ twr, const double* __restrict__ twi,
                                          int n, int L, int nfft, int fs, int es, int pad, int logw,
                                          const GlobalTile& g, double* __restrict__ power, bool first_power)
{
    const int M = L / R;
    const int nbf = n / R;
    const int tws = n / L;
    const int items = nfft * nbf;
    const bool p2 = ((M & (M - 1)) == 0) && ((nbf & (nbf - 1)) == 0);   // power-of-two stage: shifts, no divides
    const int lM = __ffs(M) - 1, lnbf = __ffs(nbf) - 1;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        int f, bf, b, n2;
        if (ZPASS) { if (p2) { f = it >> lnbf; bf = it & (nbf - 1); } else { f = it / nbf; bf = it - f * nbf; } }
        else       { bf = it >> logw; f = it & (nfft - 1); }
        if (p2) { b = bf >> lM; n2 = bf & (M - 1); } else { b = bf / M; n2 = bf - b * M; }
        const int base = b * L + n2;
        double xr[R], xi[R];
        if (IN == IO_GLOBAL) {
            const bool ok = f < g.w;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                double2 v = make_double2(0.0, 0.0);
                if (ok) v = g.base[(long long)(base + j * M) * g.row_stride + f];
                xr[j] = v.x; xi[j] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int p = base + j * M;
                const int a = f * fs + (p + (p >> pad)) * es;
                xr[j] = sre[a]; xi[j] = sim[a];
            }
        }
        Dft<R>::run(xr, xi, twr, twi, n);
        if (M > 1) {
#pragma unroll
            for (int k = 1; k < R; ++k) {
                const int ti = n2 * k * tws;
                const double wr = twr[ti], wi = twi[ti];
                const double yr = xr[k] * wr - xi[k] * wi;
                xi[k] = xr[k] * wi + xi[k] * wr;
                xr[k] = yr;
            }
        }
        if (OUT == IO_GLOBAL) {
            if (f < g.w) {
#pragma unroll
                for (int k = 0; k < R; ++k) g.base[(long long)(base + k * M) * g.row_stride + f] = make_double2(xr[k], xi[k]);
            }
        } else if (OUT == IO_POWER) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int a = (base + k * M) * es + f;
                const double v = xr[k] * xr[k] + xi[k] * xi[k];
                power[a] = first_power ? v : power[a] + v;
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int p = base + k * M;
                const int a = f * fs + (p + (p >> pad)) * es;
                sre[a] = xr[k]; sim[a] = xi[k];
            }
        }
    }
}

template <bool ZPASS, int IN, int OUT>
__device__ __forceinline__ void fft_stage_dispatch(int R, double* sre, double* sim, const double* twr, const double* twi,
                                                   int n, int L, int nfft, int fs, int es, int pad, int logw,
                                                   const GlobalTile& g, double* power, bool first_power)
{
    switch (R) {
        case 16: fft_stage<16, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 8:  fft_stage<8, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 4:  fft_stage<4, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 2:  fft_stage<2, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 3:  fft_stage<3, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 5:  fft_stage<5, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 7:  fft_stage<7, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 11: fft_stage<11, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        case 13: fft_stage<13, ZPASS, IN, OUT>(sre, sim, twr, twi, n, L, nfft, fs, es, pad, logw, g, power, first_power); break;
        default: break;
    }
}

// all stages of one axis over a tile resident in shared memory (z pass inside the splat kernel; caller syncs before).
// z plans use radices <= 8 only (2, 4, 8, 3, 5, 7): the splat kernel runs 1024 threads per SM, i.e. 64 registers per
// thread, and a radix-8 butterfly (32 registers of data) is what fits without spills.
__device__ __forceinline__ void fft_tile_z(double* sre, double* sim, const double* twr, const double* twi,
                                           const FftPlan& plan, int nfft, int fs, int pad)
{
    GlobalTile none{nullptr, 0, 0};
    int L = plan.n;
    for (int s = 0; s < plan.nstages; ++s) {
        switch (plan.radix[s]) {
            case 8: fft_stage<8, true, IO_SMEM, IO_SMEM>(sre, sim, twr, twi, plan.n, L, nfft, fs, 1, pad, 0, none, nullptr, false); break;
            case 4: fft_stage<4, true, IO_SMEM, IO_SMEM>(sre, sim, twr, twi, plan.n, L, nfft, fs, 1, pad, 0, none, nullptr, false); break;
            case 2: fft_stage<2, true, IO_SMEM, IO_SMEM>(sre, sim, twr, twi, plan.n, L, nfft, fs, 1, pad, 0, none, nullptr, false); break;
            case 3: fft_stage<3, true, IO_SMEM, IO_SMEM>(sre, sim, twr, twi, plan.n, L, nfft, fs, 1, pad, 0, none, nullptr, false); break;
            case 5: fft_stage<5, true, IO_SMEM, IO_SMEM>(sre, sim, twr, twi, plan.n, L, nfft, fs, 1, pad, 0, none, nullptr, false); break;
            case 7: fft_stage<7, true, IO_SMEM, IO_SMEM>(sre, sim, twr, twi, plan.n, L, nfft, fs, 1, pad, 0, none, nullptr, false); break;
            default: break;
        }
        L /= plan.radix[s];
        __syncthreads();
    }
}

// y/x pass over a [n][W] tile: first stage from global, last stage to global (POWER = false) or
// into the |C|^2 tile `power` (POWER = true).
template <bool POWER>
__device__ __forceinline__ void fft_tile_strided(double* sre, double* sim, const double* twr, const double* twi,
                                                 const FftPlan& plan, int W, int logw, const GlobalTile& g,
                                                 double* power, bool first_power)
{
    constexpr int LAST = POWER ? IO_POWER : IO_GLOBAL;
    const int n = plan.n;
    if (plan.nstages == 1) {
        fft_stage_dispatch<false, IO_GLOBAL, LAST>(plan.radix[0], sre, sim, twr, twi, n, n, W, 1, W, 31, logw, g, power, first_power);
        return;
    }
    int L = n;
    fft_stage_dispatch<false, IO_GLOBAL, IO_SMEM>(plan.radix[0], sre, sim, twr, twi, n, L, W, 1, W, 31, logw, g, power, first_power);
    L /= plan.radix[0];
    __syncthreads();
    for (int s = 1; s + 1 < plan.nstages; ++s) {
        fft_stage_dispatch<false, IO_SMEM, IO_SMEM>(plan.radix[s], sre, sim, twr, twi, n, L, W, 1, W, 31, logw, g, power, first_power);
        L /= plan.radix[s];
        __syncthreads();
    }
    fft_stage_dispatch<false, IO_SMEM, LAST>(plan.radix[plan.nstages - 1], sre, sim, twr, twi, n, L, W, 1, W, 31, logw, g, power, first_power);
}

__device__ __forceinline__ void load_twiddles(double* twr, double* twi, const double2* __restrict__ tw, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 w = tw[i];
        twr[i] = w.x; twi[i] = w.y;
    }
}

#ifndef MDSF_PASS_THREADS
#define MDSF_PASS_THREADS 128
#endif
#ifndef MDSF_PASS_MINBLOCKS
#define MDSF_PASS_MINBLOCKS 4
#endif
#ifndef MDSF_FAST_MINBLOCKS
#define MDSF_FAST_MINBLOCKS 5          // two-stage kernels: 5 CTAs/SM (102 registers) measured best on B200
#endif

// Pass geometry.  A y-pass CTA sits at (z tile, x, pair), an x-pass CTA at (z tile, y) and walks the pairs; element
// (row, f) of its [n][W] tile is base[row * rs + f] with base = vol + pair * ncell + outer * os + ztile * cs.
//   plain layout [x][y][z]:        y pass os = Ny*Nz, rs = Nz;     x pass os = Nz, rs = Ny*Nz;   cs = W
//   chunked layout [z/lw][x][y][lw]: y pass os = Ny*lw, rs = lw;   x pass os = lw, rs = Ny*lw;   cs = Nx*Ny*lw, W = lw
struct PassGeom {
    long long ncell, os, cs, rs;
    int nz, W;          // valid columns of tile t: min(W, nz - t*W)
};

// ------------------------------------------------------------------ y pass, in place
// grid = (z chunks, Nx, pairs); tile = [Ny][W] at fixed x; rows are W*16 contiguous bytes.
// THR x MINB: 128 x 4 for short axes ([n][8] tiles <= 36 KB); 256 x 2 / 512 x 1 keep 128-byte rows on long axes.
template <int THR, int MINB>
__global__ void __launch_bounds__(THR, MINB)
fft_y_kernel(double2* __restrict__ vol, FftPlan plan, const double2* __restrict__ tw, PassGeom pg, int logw)
{
    extern __shared__ double smem[];
    const int ny = plan.n, W = pg.W;
    double* sre = smem;
    double* sim = sre + (size_t)ny * W;
    double* twr = sim + (size_t)ny * W;
    double* twi = twr + ny;
    load_twiddles(twr, twi, tw, ny);
    __syncthreads();
    GlobalTile g;
    g.base = vol + (long long)blockIdx.z * pg.ncell + (long long)blockIdx.y * pg.os + (long long)blockIdx.x * pg.cs;
    g.row_stride = pg.rs;
    g.w = min(W, pg.nz - (int)blockIdx.x * W);
    fft_tile_strided<false>(sre, sim, twr, twi, plan, W, logw, g, nullptr, false);
}

// ------------------------------------------------------------------ x pass + |C|^2 accumulation
// grid = (z chunks, Ny); tile = [Nx][W] at fixed y; loops over the pairs of the batch, keeps
// sum_q |C_q|^2 in shared memory and adds it to the resident fp64 accumulator P with one
// read-modify-write per batch (dens.py:315-318).
template <int THR, int MINB>
__global__ void __launch_bounds__(THR, MINB)
fft_x_accum_kernel(double2* __restrict__ vol, double* __restrict__ P, FftPlan plan,
                   const double2* __restrict__ tw, PassGeom pg, int logw, int npairs)
{
    extern __shared__ double smem[];
    const int nx = plan.n, W = pg.W;
    double* sre = smem;
    double* sim = sre + (size_t)nx * W;
    double* acc = sim + (size_t)nx * W;
    double* twr = acc + (size_t)nx * W;
    double* twi = twr + nx;
    load_twiddles(twr, twi, tw, nx);
    const long long xstride = pg.rs;
    const long long off = (long long)blockIdx.y * pg.os + (long long)blockIdx.x * pg.cs;
    GlobalTile g;
    g.row_stride = xstride;
    g.w = min(W, pg.nz - (int)blockIdx.x * W);
    for (int q = 0; q < npairs; ++q) {
        __syncthreads();
        g.base = vol + (long long)q * pg.ncell + off;
        fft_tile_strided<true>(sre, sim, twr, twi, plan, W, logw, g, acc, q == 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nx * W; i += blockDim.x) {
        const int x = i >> logw, z = i & (W - 1);
        if (z < g.w) P[(long long)x * xstride + off + z] += acc[i];
    }
}

// ------------------------------------------------------------------ y pass, two-stage fast path (Ny = R1*R2)
// grid = (z chunks, Nx, pairs), one butterfly per thread and stage; in place on the volume.
template <int R1, int R2>
__global__ void __launch_bounds__(MDSF_PASS_THREADS, MDSF_FAST_MINBLOCKS)
fft_y_fast_kernel(double2* __restrict__ vol, const double2* __restrict__ tw, PassGeom pg, int logw)
{
    constexpr int NY = R1 * R2;
    extern __shared__ double smem[];
    const int W = 1 << logw;
    double* sre = smem;
    double* sim = sre + (size_t)NY * W;
    double* twr = sim + (size_t)NY * W;
    double* twi = twr + NY;
    load_twiddles(twr, twi, tw, NY);
    const int f = threadIdx.x & (W - 1), bf = threadIdx.x >> logw;
    const bool ok = (int)blockIdx.x * W + f < pg.nz;
    const long long nz = pg.rs;                    // row stride
    double2* base = vol + (long long)blockIdx.z * pg.ncell + (long long)blockIdx.y * pg.os + (long long)blockIdx.x * pg.cs + f;
    {
        double xr[R1], xi[R1];
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            double2 v = make_double2(0.0, 0.0);
            if (ok) v = base[(long long)(bf + R2 * j) * nz];
            xr[j] = v.x; xi[j] = v.y;
        }
        Dft<R1>::run(xr, xi, twr, twi, NY);
        __syncthreads();                           // twiddle table is in shared memory
#pragma unroll
        for (int k = 1; k < R1; ++k) {
            const double wr = twr[bf * k], wi = twi[bf * k];
            const double yr = xr[k] * wr - xi[k] * wi;
            xi[k] = xr[k] * wi + xi[k] * wr;
            xr[k] = yr;
        }
#pragma unroll
        for (int k = 0; k < R1; ++k) { const int a = ((k * R2 + bf) << logw) + f; sre[a] = xr[k]; sim[a] = xi[k]; }
    }
    __syncthreads();
    {
        double xr[R2], xi[R2];
#pragma unroll
        for (int j = 0; j < R2; ++j) { const int a = ((bf * R2 + j) << logw) + f; xr[j] = sre[a]; xi[j] = sim[a]; }
        Dft<R2>::run(xr, xi, twr, twi, NY);
        if (ok) {
#pragma unroll
            for (int k = 0; k < R2; ++k) base[(long long)(bf * R2 + k) * nz] = make_double2(xr[k], xi[k]);
        }
    }
}

// ------------------------------------------------------------------ x pass, two-stage, cp.async-prefetched
// Two-stage x pass (Nx = R1*R2, one butterfly per thread and stage, sum_q |C_q|^2 of its R2 outputs in registers).  The stage-1 inputs of pair q+1 are copied global -> shared
// memory with cp.async (16 bytes per copy, each thread copies exactly the R1 points it will consume, so no
// barrier guards the staging buffer) while the thread runs the butterflies of pair q: every CTA always has
// its next 32 KB in flight, which keeps the HBM queues full from few SMs (SM-partitioned pipeline) as well.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifndef MDSF_XASYNC_MINBLOCKS
#define MDSF_XASYNC_MINBLOCKS 3
#endif
template <int R1, int R2>
__global__ void __launch_bounds__(MDSF_PASS_THREADS, MDSF_XASYNC_MINBLOCKS)
fft_x_accum_async_kernel(const double2* __restrict__ vol, double* __restrict__ P, const double2* __restrict__ tw,
                         PassGeom pg, int logw, int npairs)
{
    constexpr int NX = R1 * R2;
    extern __shared__ double smem[];
    const int W = 1 << logw;
    double* sre = smem;
    double* sim = sre + (size_t)NX * W;
    double* twr = sim + (size_t)NX * W;
    double* twi = twr + NX;
    double2* stage = reinterpret_cast<double2*>(twi + NX);     // [R1][blockDim.x]: point j of thread t at stage[j*T + t]
    const int T = blockDim.x;
    load_twiddles(twr, twi, tw, NX);
    const long long xstride = pg.rs;
    const long long off = (long long)blockIdx.y * pg.os + (long long)blockIdx.x * pg.cs;
    const int f = threadIdx.x & (W - 1), bf = threadIdx.x >> logw;
    const bool ok = (int)blockIdx.x * W + f < pg.nz;
    double acc[R2];
#pragma unroll
    for (int k = 0; k < R2; ++k) acc[k] = 0.0;
    double2* mine = stage + threadIdx.x;
    if (ok) {
        const double2* base = vol + off + f;
#pragma unroll
        for (int j = 0; j < R1; ++j) cp_async16(mine + j * T, base + (long long)(bf + R2 * j) * xstride);
    }
    __syncthreads();                               // twiddle table is in shared memory
    for (int q = 0; q < npairs; ++q) {
        {
            double xr[R1], xi[R1];
            cp_async_wait_all();
#pragma unroll
            for (int j = 0; j < R1; ++j) {
                double2 v = make_double2(0.0, 0.0);
                if (ok) v = mine[j * T];
                xr[j] = v.x; xi[j] = v.y;
            }
            if (ok && q + 1 < npairs) {            // the slots are free again: fetch the next pair behind this one's math
                const double2* base = vol + (long long)(q + 1) * pg.ncell + off + f;
#pragma unroll
                for (int j = 0; j < R1; ++j) cp_async16(mine + j * T, base + (long long)(bf + R2 * j) * xstride);
            }
            Dft<R1>::run(xr, xi, twr, twi, NX);
#pragma unroll
            for (int k = 1; k < R1; ++k) {
                const double wr = twr[bf * k], wi = twi[bf * k];
                const double yr = xr[k] * wr - xi[k] * wi;
                xi[k] = xr[k] * wi + xi[k] * wr;
                xr[k] = yr;
            }
            __syncthreads();                       // previous pair's stage 2 has finished reading the tile
#pragma unroll
            for (int k = 0; k < R1; ++k) { const int a = ((k * R2 + bf) << logw) + f; sre[a] = xr[k]; sim[a] = xi[k]; }
        }
        __syncthreads();
        {
            double xr[R2], xi[R2];
#pragma unroll
            for (int j = 0; j < R2; ++j) { const int a = ((bf * R2 + j) << logw) + f; xr[j] = sre[a]; xi[j] = sim[a]; }
            Dft<R2>::run(xr, xi, twr, twi, NX);
#pragma unroll
            for (int k = 0; k < R2; ++k) acc[k] += xr[k] * xr[k] + xi[k] * xi[k];
        }
    }
    if (ok) {
#pragma unroll
        for (int k = 0; k < R2; ++k) P[(long long)(bf * R2 + k) * xstride + off + f] += acc[k];
    }
}

// ------------------------------------------------------------------ y / x pass, three radix stages (long axes)
// N = R1*R2*R3 (512 = 8*8*8, 1024 = 8*8*16, 768 = 16*16*3): same decimation-in-frequency scheme and output order as
// fft_stage (position space), written out with compile-time radices and strides.  Tile [N][W] in shared memory
// (re / im planes); every thread owns the same butterflies for every pair of the batch, so the x pass keeps
// sum_q |C_q|^2 of its stage-3 outputs in registers and touches P once (the generic kernel keeps a third [N][W]
// plane in shared memory for that and dispatches radices at run time).
//   XPASS = false: in place on the volume, grid = (z chunks, Nx, pairs)
//   XPASS = true : grid = (z chunks, Ny), loops over the pairs, P += sum |C|^2
template <int R1, int R2, int R3, int LOGW, int THR, int MINB, bool XPASS>
__global__ void __launch_bounds__(THR, MINB)
fft3_pass_kernel(double2* __restrict__ vol, double* __restrict__ P, const double2* __restrict__ tw,
                 PassGeom pg, int npairs)
{
    constexpr int N = R1 * R2 * R3, M1 = N / R1, M2 = M1 / R2;
    extern __shared__ double smem[];
    constexpr int W = 1 << LOGW, logw = LOGW;
    double* sre = smem;
    double* sim = sre + (size_t)N * W;
    double* twr = sim + (size_t)N * W;
    double* twi = twr + N;
    load_twiddles(twr, twi, tw, N);
    const int z0 = blockIdx.x * W, nz = pg.nz;
    // XPASS: rows are x, the CTA sits at y = blockIdx.y and walks the pairs; else rows are y at x = blockIdx.y of pair blockIdx.z
    const long long row = pg.rs;
    const long long pair_stride = pg.ncell;
    const long long off = (long long)blockIdx.y * pg.os + (long long)blockIdx.x * pg.cs + (XPASS ? 0LL : (long long)blockIdx.z * pg.ncell);
    constexpr int NB1 = M1, NB2 = N / R2, NB3 = N / R3;      // butterflies per column and stage
    constexpr int items1 = NB1 * W, items2 = NB2 * W, items3 = NB3 * W;
    constexpr int MAXI3 = (items3 + THR - 1) / THR;           // stage-3 butterflies per thread
    double acc[MAXI3][R3];
    if (XPASS) {
#pragma unroll
        for (int i = 0; i < MAXI3; ++i)
#pragma unroll
            for (int k = 0; k < R3; ++k) acc[i][k] = 0.0;
    }
    const int nq = XPASS ? npairs : 1;
    for (int q = 0; q < nq; ++q) {
        double2* base = vol + off + (long long)q * pair_stride;
        __syncthreads();                                       // twiddles loaded / previous pair's stage 3 done
        // ---- stage 1: points n2 + M1*j straight from global memory
        for (int it = threadIdx.x; it < items1; it += THR) {
            const int f = it & (W - 1), n2 = it >> logw;
            const bool ok = z0 + f < nz;
            double xr[R1], xi[R1];
#pragma unroll
            for (int j = 0; j < R1; ++j) {
                double2 v = make_double2(0.0, 0.0);
                if (ok) v = base[(long long)(n2 + M1 * j) * row + f];
                xr[j] = v.x; xi[j] = v.y;
            }
            Dft<R1>::run(xr, xi, twr, twi, N);
#pragma unroll
            for (int k = 1; k < R1; ++k) {
                const double wr = twr[n2 * k], wi = twi[n2 * k];
                const double yr = xr[k] * wr - xi[k] * wi;
                xi[k] = xr[k] * wi + xi[k] * wr;
                xr[k] = yr;
            }
#pragma unroll
            for (int k = 0; k < R1; ++k) { const int a = ((n2 + M1 * k) << logw) + f; sre[a] = xr[k]; sim[a] = xi[k]; }
        }
        __syncthreads();
        // ---- stage 2: block b of M1 points, points b*M1 + n2 + M2*j, twiddle w_N^(R1 n2 k)
        for (int it = threadIdx.x; it < items2; it += THR) {
            const int f = it & (W - 1), bf = it >> logw;
            const int b = bf / M2, n2 = bf - b * M2;
            const int p0 = b * M1 + n2;
            double xr[R2], xi[R2];
#pragma unroll
            for (int j = 0; j < R2; ++j) { const int a = ((p0 + M2 * j) << logw) + f; xr[j] = sre[a]; xi[j] = sim[a]; }
            Dft<R2>::run(xr, xi, twr, twi, N);
#pragma unroll
            for (int k = 1; k < R2; ++k) {
                const double wr = twr[n2 * k * R1], wi = twi[n2 * k * R1];
                const double yr = xr[k] * wr - xi[k] * wi;
                xi[k] = xr[k] * wi + xi[k] * wr;
                xr[k] = yr;
            }
#pragma unroll
            for (int k = 0; k < R2; ++k) { const int a = ((p0 + M2 * k) << logw) + f; sre[a] = xr[k]; sim[a] = xi[k]; }
        }
        __syncthreads();
        // ---- stage 3: R3 consecutive points of block bf; outputs stay at bf*R3 + k
#pragma unroll
        for (int i = 0; i < MAXI3; ++i) {
            const int it = threadIdx.x + i * THR;
            if (it < items3) {
                const int f = it & (W - 1), bf = it >> logw;
                double xr[R3], xi[R3];
#pragma unroll
                for (int j = 0; j < R3; ++j) { const int a = ((bf * R3 + j) << logw) + f; xr[j] = sre[a]; xi[j] = sim[a]; }
                Dft<R3>::run(xr, xi, twr, twi, N);
                if (XPASS) {
#pragma unroll
                    for (int k = 0; k < R3; ++k) acc[i][k] += xr[k] * xr[k] + xi[k] * xi[k];
                } else if (z0 + f < nz) {
#pragma unroll
                    for (int k = 0; k < R3; ++k) base[(long long)(bf * R3 + k) * row + f] = make_double2(xr[k], xi[k]);
                }
            }
        }
    }
    if (XPASS) {
#pragma unroll
        for (int i = 0; i < MAXI3; ++i) {
            const int it = threadIdx.x + i * THR;
            if (it < items3) {
                const int f = it & (W - 1), bf = it >> logw;
                if (z0 + f < nz) {
#pragma unroll
                    for (int k = 0; k < R3; ++k) P[(long long)(bf * R3 + k) * row + off + f] += acc[i][k];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ library-FFT path helper
// P += sum_q |vol_q|^2 after a cuFFT Z2Z (grids whose sizes have prime factors > 13)
static __global__ void accumulate_power_kernel(const double2* __restrict__ vol, double* __restrict__ P,
                                        long long ncell, int npairs)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell;
         i += (long long)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int q = 0; q < npairs; ++q) {
            const double2 v = vol[(long long)q * ncell + i];
            a += v.x * v.x + v.y * v.y;
        }
        P[i] += a;
    }
}

// ------------------------------------------------------------------ S(q) read-out
// sf[kx][ky][kz] = (P[pos(k)] + P[pos(-k)]) / 2 for kz in [0, Nz/2]: the half spectrum that
// dens.py:318 accumulates.  rev* map a frequency index to its position (identity for cuFFT).
static __global__ void export_sf_kernel(const double* __restrict__ P, double* __restrict__ sf,
                                 const int* __restrict__ revx, const int* __restrict__ revy,
                                 const int* __restrict__ revz, GridParams gp)
{
    const int nx = gp.n[0], ny = gp.n[1], nz = gp.n[2];
    const int m = nz / 2 + 1;
    const long long total = (long long)nx * ny * m;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int kz = (int)(i % m);
        const long long t = i / m;
        const int ky = (int)(t % ny), kx = (int)(t / ny);
        const long long a = vol_index(gp, revx[kx], revy[ky], revz[kz]);
        const long long b = vol_index(gp, revx[(nx - kx) % nx], revy[(ny - ky) % ny], revz[(nz - kz) % nz]);
        sf[i] = 0.5 * (P[a] + P[b]);
    }
}

// un-pack a pair volume (before any FFT) into the real density of one frame (debug tap)
static __global__ void unpack_density_kernel(const double2* __restrict__ vol, double* __restrict__ d1,
                                      long long ncell, int part)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell;
         i += (long long)gridDim.x * blockDim.x) d1[i] = part ? vol[i].y : vol[i].x;
}
