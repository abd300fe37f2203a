// Shared definitions of the mdsf engine (sm_100a).  See include/mdsf.h for the C ABI and
// DESIGN.md for the data layout.  Reference line numbers refer to joeyelk/MD-Structure-Factor.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MDSF_MAX_BATCH 64          // frames per device batch (scale factors travel as kernel params)
#define MDSF_MAX_TILE_COLS 32      // columns per splat tile (one bit each in a 32-bit hit mask)
#define MDSF_ATOM_BITS 26          // atom index bits inside a pair payload
#define MDSF_MAX_ATOMS (1 << MDSF_ATOM_BITS)
#define MDSF_MAX_RADIX_STAGES 12

// Frame-invariant geometry, passed to kernels by value.
struct GridParams {
    int n[3];           // Nspatialgrid (dens.py:181-189)
    int nb;             // Nborder (dens.py:231)
    int fold_mode;      // 0 = reference corner rule (dens.py:107), 1 = periodic
    int separable;      // ucell couples z to nothing else -> exp splits into xy and z factors
    int tx, ty;         // splat tile, in (x,y) columns; a tile spans all z
    int ntx, nty;       // tiles per dimension
    int natoms;
    int debug_skip;     // profiling aid (MDSF_SPLAT_SKIP): 1 phase B, 2 z FFT, 4 table staging, 8 store, 16 whole list loop
    int nslab, zs;      // a tile column is owned in nslab z slabs of zs cells (nslab * tx*ty = 128 owner threads)
    int nzp;            // padded z length of one column in shared memory
    int pad_shift;      // column position p is stored at p + (p >> pad_shift)
    double dr[3];       // dens.py:202
    double box[3];      // mean box (dens.py:52)
    double u[9];        // ucell row-major (dens.py:301)
    // separable case: |c|^2 = cxx bx^2 + cyy by^2 + 2 gxy bx by + czz bz^2
    double cxx, cyy, gxy, czz;
    long long tstride;  // doubles of per-atom factor tables per frame
    double fx_scale, fx_inv;   // fixed-point scale 2^(52-e) and its inverse (tile-atomic / scatter modes)
};

struct TypeTable {
    const double* amp;       // Nel / sigma^3
    const double* two_sig2;  // 2 sigma^2
    const int*    halfw;     // [ntypes][3]
    const double* ctab;      // per type [2Ax][2Ay]: exp(-2 gxy dx dy i j / (2 sigma^2)); nullptr when gxy == 0
    const int*    ctab_off;  // [ntypes] offsets into ctab
    const unsigned* toff;    // [natoms] offset of an atom's factor tables inside a frame block
};

// One rescaled+wrapped atom of one frame: what dens.py:285-287 derives per atom.
struct AtomRec {
    double r[3];   // coordinate as float64 (value of the coords-dtype number)
    int ir[3];     // trunc(r/dr)  (dens.py:285)
    int type;
    unsigned tbase;   // offset of this atom's factor tables (frame block included), in doubles
    unsigned pad_;
};

struct BatchScales {
    double a[MDSF_MAX_BATCH][3];   // avgdims/dims per frame (dens.py:53)
};

// The stamp of an atom covers padded-grid indices p in [ir-A, ir+A) (dens.py:292-297, with the
// +Nborder offset removed).  Split by side s in {-1,0,+1} (low padding / cell / high padding);
// remap_grid_tcl (dens.py:86-108) moves side s to c = p - s*N.
__host__ __device__ inline void stamp_segment(int ir, int A, int N, int s, int& plo, int& phi) {
    int p0 = ir - A, p1 = ir + A;
    if (s < 0)       { plo = p0;               phi = p1 < 0 ? p1 : 0; }
    else if (s == 0) { plo = p0 > 0 ? p0 : 0;  phi = p1 < N ? p1 : N; }
    else             { plo = p0 > N ? p0 : N;  phi = p1; }
}

// number of tiles (width t) a stamp touches along one dimension, summed over its segments
__host__ __device__ inline int stamp_tiles_1d(int ir, int A, int N, int t) {
    int cnt = 0;
    for (int s = -1; s <= 1; ++s) {
        int plo, phi;
        stamp_segment(ir, A, N, s, plo, phi);
        if (phi > plo) {
            int clo = plo - s * N, chi = phi - s * N;
            cnt += (chi - 1) / t - clo / t + 1;
        }
    }
    return cnt;
}

// z slabs (bit s = slab s of zs cells) the image (sx,sy) of an atom touches after the fold.
// The three z segments of the stamp (low padding / cell / high padding) move by shlo / 0 / shhi;
// in the 8 corner regions the reference picks the z block by the y side (dens.py:107), which
// turns the shift into +-Nborder when the y side differs from the z side.
__host__ __device__ inline unsigned image_slabmask(int irz, int Az, int sx, int sy, int nz, int nb, int fold_mode,
                                                   int zs, int& shlo, int& shhi, int& kA, int& kB) {
    const int pz0 = irz - Az, nzr = 2 * Az;
    kA = -pz0 < 0 ? 0 : (-pz0 > nzr ? nzr : -pz0);
    kB = nz - pz0 < 0 ? 0 : (nz - pz0 > nzr ? nzr : nz - pz0);
    const bool corner = (sx != 0 && sy != 0 && fold_mode == 0);
    shlo = (corner && sy != -1) ? nb : nz;
    shhi = (corner && sy != 1) ? -nb : -nz;
    unsigned m = 0;
    if (kA > 0)   for (int sl = (pz0 + shlo) / zs; sl <= (pz0 + kA - 1 + shlo) / zs; ++sl) m |= 1u << sl;
    if (kB > kA)  for (int sl = (pz0 + kA) / zs; sl <= (pz0 + kB - 1) / zs; ++sl) m |= 1u << sl;
    if (nzr > kB) for (int sl = (pz0 + kB + shhi) / zs; sl <= (pz0 + nzr - 1 + shhi) / zs; ++sl) m |= 1u << sl;
    return m;
}
