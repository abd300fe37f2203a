// Shared definitions of the mdsf engine (sm_100a).  See include/mdsf.h for the C ABI and
// DESIGN.md for the data layout.  Reference line numbers refer to joeyelk/MD-Structure-Factor.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MDSF_MAX_BATCH 64          // frames per device batch (scale factors travel as kernel params)
#define MDSF_MAX_ATOMS (1 << 28)
#define MDSF_MAX_RADIX_STAGES 12
#define MDSF_MAX_STAMP 1023        // 2*A: stamp indices travel in 10-bit fields / int offsets
#ifndef MDSF_SPLAT_WARPS
#define MDSF_SPLAT_WARPS 16        // warps per splat CTA: 2 CTAs x 512 threads at 64 registers.  With the interleaved tile the z stages no longer
                                   // spill at 64 registers, 32 work items are exactly two per warp and 512 butterflies per stage one per thread
                                   // (c3 splat 19.1 -> 17.2 ms against 12 warps at 80 registers; c1 / c4 lose 3-4 %)
#endif

// Frame-invariant geometry, passed to kernels by value.
struct GridParams {
    int n[3];           // Nspatialgrid (dens.py:181-189)
    int nb;             // Nborder (dens.py:231)
    int fold_mode;      // 0 = reference corner rule (dens.py:107), 1 = periodic
    int separable;      // ucell couples z to nothing else -> exp splits into xy and z factors
    int lcol;           // splat tile = 2^lcol (x,y) columns over all z: TX = 2^((lcol+1)/2), TY = 2^(lcol/2)
    int ntx, nty;       // tiles per dimension
    int nslab, zw;      // z slabs per column, slab width zw = (256 >> lcol) / sub cells: one list per slab and tile
    int sub;            // lists a splat warp walks side by side (1 or 2 half-warp groups)
    int zlane;          // 1: z-lane splat variant (sub = 1: lanes walk z, a lane holds up to 8 columns; mdsf_splat.cuh)
    int natoms;
    int nzp;            // padded z length of one column in shared memory
    int pad_shift;      // column position p is stored at p + (p >> pad_shift) ...
    int zilv;           // 1: interleaved tile [col][nz + 1] of 16-byte complex cells (compile-time z path; mdsf_splat.cuh tile_index)
    // volume layout: element (x, y, z) of a pair volume sits at ((z / lw * Nx + x) * Ny + y) * lw + z % lw.
    // lw = Nz is the plain C order [x][y][z]; a smaller lw makes every (x, z-chunk) row block of the y pass one
    // contiguous run and lets the y -> x hand-over of a chunk stay L2-resident.
    int lw, nch;        // chunk width, chunks per column (nch * lw = Nz)
    double dr[3];       // dens.py:202
    double box[3];      // mean box (dens.py:52)
    double u[9];        // ucell row-major (dens.py:301)
    // separable case: |c|^2 = cxx bx^2 + cyy by^2 + 2 gxy bx by + czz bz^2
    double cxx, cyy, gxy, czz;
    long long tstride;  // doubles of per-atom factor tables per frame
    double fx_scale, fx_inv;   // fixed-point scale 2^(52-e) and its inverse (amax < 2^e)
};

__host__ __device__ inline long long vol_index(const GridParams& gp, int x, int y, int z) {
    const int ch = z / gp.lw, zw = z - ch * gp.lw;
    return (((long long)ch * gp.n[0] + x) * gp.n[1] + y) * gp.lw + zw;
}

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
    unsigned pad_;    // 1: K1 rejected the atom (stamp outside the padded grid / NaN)
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

// Fold shift of the z segment `sz` of image (sx, sy): destination z = padded z + shift.  Faces and edges move by
// -sz*Nz; in the 8 corner regions the reference picks the z block by the y side (dens.py:107), which turns the
// shift into +-Nborder when the y side differs from the z side.
__host__ __device__ inline int fold_shift_z(int sx, int sy, int sz, int nz, int nb, int fold_mode) {
    if (sz == 0) return 0;
    const bool corner = (sx != 0 && sy != 0 && fold_mode == 0);
    if (sz < 0) return (corner && sy != -1) ? nb : nz;
    return (corner && sy != 1) ? -nb : -nz;
}

// number of width-t bins the folded stamp [ir-A, ir+A) touches along one dimension, summed over its segments;
// `sh_lo` / `sh_hi` are the shifts of the low / high padding segment (normally +N / -N)
__host__ __device__ inline int stamp_bins_1d(int ir, int A, int N, int t, int sh_lo, int sh_hi) {
    int cnt = 0;
    for (int s = -1; s <= 1; ++s) {
        int plo, phi;
        stamp_segment(ir, A, N, s, plo, phi);
        if (phi > plo) {
            const int sh = s < 0 ? sh_lo : (s > 0 ? sh_hi : 0);
            cnt += (phi - 1 + sh) / t - (plo + sh) / t + 1;
        }
    }
    return cnt;
}

// One (atom image, tile, z slab) pair, prepared by the binning kernel; what a splat warp consumes.
//   separable ucell:  x = index of EZ[k] for slab-local z 0 (tables[x + zz], zz in [zoff, zoff + zlen))
//                     y = index of EX[i] for tile column x 0, z = index of EY[j] for tile column y 0 (columns outside
//                     the stamp read the zero pads of the table block, see table_doubles in mdsf_prep.cuh)
//                     w = zoff | zlen << 7
//   general ucell:    x, y, z = padded-grid index of slab-local z 0 / tile column x 0 / tile column y 0
//                     w = cx0 | cx1 << 3 | cy0 << 7 | cy1 << 10 | zoff << 14 | zend << 21   (clip box, [lo, hi))
// PairAux (monoclinic or general ucell only): x = index of the cross-term table entry of tile column (0,0) / atom index,
//                                             y = row stride 2*Ay of that table / unused
typedef uint4 PairRec;
typedef uint2 PairAux;
