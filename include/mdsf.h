/*
 * mdsf.h -- C ABI of libmdsf.so, the B200 (sm_100a) engine behind dens.compute_sf().
 *
 * The reference (joeyelk/MD-Structure-Factor) is pure Python and has no FFI; its whole hot
 * path is the frame loop of dens.compute_sf (reference dens.py:277-321).  This header is the
 * boundary a maintainer binds with ctypes (see INTEGRATION.md for the stub) to replace exactly
 * that loop.  Every entry point cites the reference lines whose work it takes over.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative MDSF_E* code on failure; mdsf_last_error() gives the message for the calling
 * thread's last failure.  Host arrays stay owned by the caller.  One handle = one GPU + its
 * streams; calls on one handle are not thread-safe, different handles are independent.
 * There is no CPU fallback: without a CUDA device mdsf_create() fails with MDSF_ECUDA.
 */
#ifndef MDSF_H
#define MDSF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDSF_ABI_VERSION 1

enum {
    MDSF_OK = 0,
    MDSF_EINVAL = -1,   /* bad argument / unsupported configuration            */
    MDSF_ECUDA = -2,    /* CUDA runtime / cuFFT failure (no device, OOM, ...)  */
    MDSF_ERANGE = -3,   /* an atom's stamp leaves the padded grid (the reference
                           raises a numpy shape error or corrupts memory here)  */
    MDSF_ESTATE = -4    /* call order violated (e.g. push before set_atoms)     */
};

enum { MDSF_F32 = 0, MDSF_F64 = 1 };
enum { MDSF_FOLD_REFERENCE = 0,   /* reproduce the corner rule of dens.py:107 (default) */
       MDSF_FOLD_PERIODIC = 1 };  /* mathematically periodic fold                       */
enum { MDSF_FFT_AUTO = 0, MDSF_FFT_NATIVE = 1, MDSF_FFT_CUFFT = 2 };

typedef struct mdsf_handle mdsf_handle;

/* Everything that is frame-invariant.  The host computes N, dr, box, half widths and B with
 * the same numpy expressions as dens.py:179-231 and passes them in, so they are identical to
 * the reference's by construction. */
typedef struct mdsf_config {
    int32_t abi_version;     /* MDSF_ABI_VERSION                                              */
    int32_t device;          /* CUDA device ordinal                                           */
    int32_t n[3];            /* Nspatialgrid, even (dens.py:181-189)                          */
    int32_t nborder;         /* Nborder B (dens.py:231)                                       */
    double  dr[3];           /* L / Nspatialgrid, float64 (dens.py:202)                       */
    double  box[3];          /* mean box L, exact value of the dims dtype (dens.py:52)        */
    double  ucell[9];        /* row-major 3x3, rows = unit lattice vectors (dens.py:301)      */
    int32_t ntypes;          /* number of distinct labels in typ                              */
    const double*  amp;      /* [ntypes]  Nel / sigma^3            (dens.py:308)              */
    const double*  two_sig2; /* [ntypes]  2 * sigma^2              (dens.py:308)              */
    const int32_t* halfw;    /* [ntypes][3] trunc(get_borders)     (dens.py:38-43,287)        */
    int32_t coord_dtype;     /* MDSF_F32 / MDSF_F64: dtype of the coordinate array            */
    int32_t arith_dtype;     /* dtype numpy promotes (coords, dims) to for rescale and wrap   */
    int32_t fold_mode;       /* MDSF_FOLD_*                                                   */
    int32_t fft_mode;        /* MDSF_FFT_*                                                    */
    int32_t batch_frames;    /* frames per device batch (even, >= 2); 0 = pick automatically  */
    int32_t tile_x, tile_y;  /* splat tile in columns (2x2, 4x2, 4x4 or 8x4); 0 = pick automatically */
    int32_t keep_density;    /* keep per-frame densities of the last batch for the debug tap  */
    int32_t splat_mode;      /* must be 0 (round 1 had selectable splat variants; one remains) */
    int32_t reserved[6];
} mdsf_config;

/* Create / destroy an engine.  Replaces the allocations of dens.py:237-263. */
int mdsf_create(const mdsf_config* cfg, mdsf_handle** out);
int mdsf_destroy(mdsf_handle* h);

/* Atom types are frame-invariant (dens.py:287,305-306 look typ[im] up per atom per frame). */
int mdsf_set_atoms(mdsf_handle* h, int64_t natoms, const int32_t* type_id);

/* Pinned host memory for trajectory staging (new; the reference holds the trajectory in
 * pageable numpy memory, main_gromacs.py:200-202). */
int mdsf_host_alloc(size_t bytes, void** out);
int mdsf_host_free(void* p);
int mdsf_host_register(void* p, size_t bytes);
int mdsf_host_unregister(void* p);

/* Queue `nframes` consecutive frames (coords[nframes][natoms][3], coord_dtype) for the whole
 * per-frame path: rescale (dens.py:56-58), PBC wrap (dens.py:209-221), cell index
 * (dens.py:285), Gaussian stamp (dens.py:287-308), fold (dens.py:311), FFT (dens.py:313),
 * |F|^2 accumulation (dens.py:315-318).  `scale` is a[it][0..2] = avgdims/dims per frame as
 * doubles holding the exact dims-dtype values.  Atoms with index in [wrap_lo, wrap_hi) are
 * wrapped (the reference's >=1e6-atom branch only wraps the first nframes atoms).
 * If `write_back` is non-zero the rescaled+wrapped coordinates are copied back into `coords`
 * (the reference mutates its argument in place); they are valid after mdsf_sync().
 * Asynchronous: returns once the work is queued; `coords` must stay valid and unmodified
 * until mdsf_sync() (pinned memory) -- pageable memory is staged before return.  `coords` may
 * also be a DEVICE pointer (frames already resident in HBM): the copy is then device-to-device. */
int mdsf_push_frames(mdsf_handle* h, void* coords, int64_t nframes, const double* scale,
                     int64_t wrap_lo, int64_t wrap_hi, int32_t write_back);

/* Streaming loaders (load_traj.NpzFrameStream -> dens.compute_sf_stream) reuse their pinned chunk buffers:
 * mdsf_input_mark() returns a ticket behind every host<->device copy queued so far, mdsf_input_wait() blocks
 * until those copies are done (the frames themselves may still be in the splat / FFT kernels). */
int mdsf_input_mark(mdsf_handle* h, int64_t* ticket);
int mdsf_input_wait(mdsf_handle* h, int64_t ticket);

/* RANDOM_NOISE mode of the reference (dens.py:279-280): feed ready-made real densities
 * d1[nframes][Nx][Ny][Nz] (float64, host) straight into FFT + accumulation. */
int mdsf_push_density(mdsf_handle* h, const double* d1, int64_t nframes);

/* Wait for all queued frames; reports deferred device-side errors (MDSF_ERANGE). */
int mdsf_sync(mdsf_handle* h);

/* sf[Nx][Ny][Nz/2+1] (float64), the sum over all pushed frames of |rfftn(d1)|^2, i.e. the
 * array dens.py:318 accumulates.  Host or device destination. */
int mdsf_read_sf(mdsf_handle* h, double* sf_host);
int mdsf_export_sf_device(mdsf_handle* h, void* sf_device);
/* Plot grids of dens.py:323-344 assembled on the GPU from the resident S(q): sfplt = get_dplot(sf) (dens.py:142-163)
 * [Nx-2][Ny-2][2M-3], kgrid [Nx][Ny][M][4] (dens.py:325-334) and kgridplt [Nx-2][Ny-2][2M-3][4] (dens.py:336-344), M = Nz/2+1,
 * written to caller-owned host arrays (any of the three may be NULL).  kx[Nx], ky[Ny], kz[M] and px[Nx-2], py[Ny-2],
 * pz[2M-3] are the per-index axis values of the reference's expressions, evaluated by the caller; the GPU broadcasts,
 * gathers and interleaves only, so all values are bit-identical to the numpy route. */
int mdsf_export_plot_grids(mdsf_handle* h, const double* kx, const double* ky, const double* kz,
                           const double* px, const double* py, const double* pz,
                           double* sfplt_host, double* kgrid_host, double* kgridplt_host);
/* Zero the accumulator (start a new trajectory on the same handle). */
/* Cylindrical average of plot2d.PLOT_RAD_NEW (reference plot2d.py:575-632), stand-alone (no handle): sf_host[n0][n1][n2]
 * and its axis vectors X, Y, Z are channel 3 and the coordinate channels of kgridplt (plot2d.py:571-574); binv9 = inverse
 * of the reciprocal-vector matrix (plot2d.py:594-598, row-major); rarr[rbins], ntheta[rbins] (theta samples per ring,
 * plot2d.py:616), zar[zbins].  oa_host[rbins][zbins] = mean over theta of the trilinear interpolant (scipy
 * RegularGridInterpolator semantics, NaN outside the grid). */
int mdsf_cylindrical_average(int device, const double* sf_host, int n0, int n1, int n2, const double* X, const double* Y,
                             const double* Z, const double* binv9, int rbins, const double* rarr, const int* ntheta,
                             int zbins, const double* zar, double* oa_host);
int mdsf_reset(mdsf_handle* h);

/* Parity taps (tests only; each synchronises). `frame` indexes the frames of the LAST push. */
int mdsf_debug_cell_indices(mdsf_handle* h, int64_t frame, int32_t* ir_out /* [natoms][3] */);
int mdsf_debug_coords(mdsf_handle* h, int64_t frame, double* r_out /* [natoms][3] */);
int mdsf_debug_density(mdsf_handle* h, int64_t frame, double* d1_out /* [Nx][Ny][Nz] */);

/* Introspection for bench.py / DESIGN.md: kernel launches issued so far, frames processed,
 * name of the FFT path in use ("native" / "cufft"), device time of the last batch. */
int64_t mdsf_kernel_launches(const mdsf_handle* h);
int64_t mdsf_frames_done(const mdsf_handle* h);
const char* mdsf_fft_path(const mdsf_handle* h);
const char* mdsf_splat_path(const mdsf_handle* h);   /* "register-ortho" / "register-mono" / "register-general" */
int mdsf_batch_frames(const mdsf_handle* h);
/* Coordinate pre-transform of the CLI (reference main_gromacs.py:204-207), applied by the first kernel to every
 * frame pushed afterwards, before the rescale: y <- y / sin_theta; x <- x - y * cos_theta, evaluated in float64 and
 * rounded to the coordinate dtype after each line exactly as numpy does for `T[..., 1] / np.sin(theta)`.
 * With write_back the transformed (and rescaled, wrapped) coordinates are what comes back. */
int mdsf_set_pretransform(mdsf_handle* h, int32_t enabled, double sin_theta, double cos_theta);
/* Launch geometry for bench.py / DESIGN.md: out[0..7] = splat tile columns in x, in y, z-slab width of one warp,
 * slabs per column, volume layout chunk width (Nz = plain [x][y][z]), y-pass tile width, x-pass tile width,
 * KiB of shared memory per splat CTA. */
int mdsf_geometry(const mdsf_handle* h, int32_t* out8);
/* Record CUDA events around every stage of subsequent batches; query the accumulated
 * per-stage device milliseconds: out[0..5] = copy, prep+bin, splat+zfft, y pass, x pass+accumulate
 * (library path: cuFFT, accumulate), compute-stream total */
int mdsf_enable_timing(mdsf_handle* h, int32_t on);
int mdsf_stage_ms(mdsf_handle* h, double* out6, int64_t* batches);
/* CUDA-event stopwatch around everything queued between the two calls (start: copy stream,
 * stop: compute stream; stop waits for completion). */
int mdsf_timer_start(mdsf_handle* h);
int mdsf_timer_stop(mdsf_handle* h, double* ms_out);

const char* mdsf_last_error(void);
int mdsf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDSF_H */
