// micro-benchmark: throughput of scattered 64-bit integer / fp64 reductions into an L2-resident buffer
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__global__ void k(unsigned long long* buf, double* dbuf, size_t ncell, int nz, long long nstamps, int run) {
    // each thread = one (atom, column): `run` consecutive z cells at a random (column, z0)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nstamps; i += (long long)gridDim.x * blockDim.x) {
        unsigned h = hash((unsigned)i * 2654435761u + 12345u);
        size_t base = ((size_t)h % (ncell / nz)) * nz + (hash(h) % (nz - run));
        for (int kk = 0; kk < run; ++kk) {
            if (MODE == 0) atomicAdd(&buf[base + kk], (unsigned long long)(h & 1023) + kk);
            else if (MODE == 1) atomicAdd(&dbuf[base + kk], 1.0 + kk);
            else if (MODE == 2) { asm volatile("red.global.add.u64 [%0], %1;" ::"l"(&buf[base + kk]), "l"((unsigned long long)(h & 1023) + kk) : "memory"); }
        }
    }
}
__global__ void smem_atom(unsigned long long* out, int iters) {
    __shared__ unsigned long long s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = 0;
    __syncthreads();
    unsigned h = hash(threadIdx.x + blockIdx.x * 977);
    for (int it = 0; it < iters; ++it) { h = hash(h); atomicAdd(&s[h & 4095], (unsigned long long)it); }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[0] + s[17];
}
int main() {
    const int nz = 256;
    for (size_t mb : {16, 64, 268}) {
        size_t ncell = mb * 1024 * 1024 / 8;
        unsigned long long* buf; cudaMalloc(&buf, ncell * 8); cudaMemset(buf, 0, ncell * 8);
        long long nst = 1700000LL * 16;   // (atom, column) visits of 16 c2 frames
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int mode = 0; mode < 3; ++mode) for (int run : {4, 1}) {
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) k<0><<<148 * 16, 256>>>(buf, (double*)buf, ncell, nz, nst, run);
                if (mode == 1) k<1><<<148 * 16, 256>>>(buf, (double*)buf, ncell, nz, nst, run);
                if (mode == 2) k<2><<<148 * 16, 256>>>(buf, (double*)buf, ncell, nz, nst, run);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
            }
            printf("buf %4zu MB mode %d (0 u64 atomicAdd, 1 f64 atomicAdd, 2 red.u64) run %d: %.3f ms -> %.1f G adds/s  (%.1f us per c2 frame)\n",
                   mb, mode, run, best, nst * run / best / 1e6, best * 1e3 / 16 * (4.0 / run));
        }
        cudaFree(buf);
    }
    unsigned long long* o; cudaMalloc(&o, 148 * 8 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); smem_atom<<<148 * 4, 256>>>(o, 4096); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("smem u64 atomicAdd: %.3f ms for %lld adds -> %.1f G adds/s\n", ms, 148LL * 4 * 256 * 4096, 148.0 * 4 * 256 * 4096 / ms / 1e6);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
