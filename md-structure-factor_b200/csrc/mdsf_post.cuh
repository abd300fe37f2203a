// Post-processing on the GPU: the cylindrical average of plot2d.PLOT_RAD_NEW (reference
// plot2d.py:575-632) -- for every (r, z) the mean over theta of the trilinearly interpolated structure factor at
// (r cos t, r sin t, z) . b_inv.  The reference evaluates scipy's RegularGridInterpolator on one ring of points per r in a
// Python loop (400 rings, tens of seconds at c1); here one thread owns one (r, z) output and walks its ring.
// Interpolation follows scipy's linear method step by step: interval i = searchsorted(grid, x) - 1 clipped to [0, n-2],
// t = (x - grid[i]) / (grid[i+1] - grid[i]), the eight corner terms added in itertools.product order with the weight
// formed as ((1 * wx) * wy) * wz, NaN outside the grid (bounds_error=False); the ring sum runs in theta order like
// np.average over that axis.  Only cos / sin differ from numpy in the last place.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ int post_interval(const double* __restrict__ g, int n, double x) {
    int lo = 0, hi = n;                        // np.searchsorted(g, x), side = 'left': first index with g[idx] >= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] < x) lo = mid + 1; else hi = mid;
    }
    int i = lo - 1;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    return i;
}

static __global__ void cyl_average_kernel(const double* __restrict__ sf, int n0, int n1, int n2,
                                          const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ Z,
                                          const double* __restrict__ binv /* [9] row-major */, int rbins, const double* __restrict__ rarr,
                                          const int* __restrict__ ntheta, int zbins, const double* __restrict__ zar,
                                          double* __restrict__ oa /* [rbins][zbins] */)
{
    const long long total = (long long)rbins * zbins;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(e % zbins), ir = (int)(e / zbins);
        const double r = rarr[ir], z = zar[iz];
        const int nt = ntheta[ir];
        double sum = 0.0;
        for (int it = 0; it < nt; ++it) {
            // thetas = np.linspace(0, 2 pi, nt, endpoint=False): it * (2 pi / nt)
            const double t = (double)it * (6.283185307179586 / (double)nt);
            const double px = r * cos(t), py = r * sin(t);
            // pts @ b_inv (row vector times matrix)
            const double q0 = px * binv[0] + py * binv[3] + z * binv[6];
            const double q1 = px * binv[1] + py * binv[4] + z * binv[7];
            const double q2 = px * binv[2] + py * binv[5] + z * binv[8];
            double v;
            if (q0 < X[0] || q0 > X[n0 - 1] || q1 < Y[0] || q1 > Y[n1 - 1] || q2 < Z[0] || q2 > Z[n2 - 1] ||
                !(q0 == q0) || !(q1 == q1) || !(q2 == q2)) {
                v = nan;
            } else {
                const int i0 = post_interval(X, n0, q0), i1 = post_interval(Y, n1, q1), i2 = post_interval(Z, n2, q2);
                const double t0 = (q0 - X[i0]) / (X[i0 + 1] - X[i0]);
                const double t1 = (q1 - Y[i1]) / (Y[i1 + 1] - Y[i1]);
                const double t2 = (q2 - Z[i2]) / (Z[i2 + 1] - Z[i2]);
                const double w0[2] = {1.0 - t0, t0}, w1[2] = {1.0 - t1, t1}, w2[2] = {1.0 - t2, t2};
                v = 0.0;
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const double w = __dmul_rn(__dmul_rn(__dmul_rn(1.0, w0[a]), w1[b]), w2[c]);
                            v = __dadd_rn(v, __dmul_rn(sf[((long long)(i0 + a) * n1 + (i1 + b)) * n2 + (i2 + c)], w));
                        }
            }
            sum = __dadd_rn(sum, v);
        }
        oa[e] = sum / (double)nt;
    }
}
