"""ctypes binding of libmdsf.so (C ABI in include/mdsf.h).

This is the only place Python touches the CUDA engine.  There is no CPU fallback: if the shared
library is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDSF_LIB") or os.path.join(_HERE, "libmdsf.so")      # MDSF_LIB: A/B builds of the same ABI
ABI_VERSION = 1

F32, F64 = 0, 1
FOLD_REFERENCE, FOLD_PERIODIC = 0, 1
FFT_AUTO, FFT_NATIVE, FFT_CUFFT = 0, 1, 2

EXPORTS = [
    "mdsf_create", "mdsf_destroy", "mdsf_set_atoms", "mdsf_host_alloc", "mdsf_host_free",
    "mdsf_host_register", "mdsf_host_unregister", "mdsf_push_frames", "mdsf_push_density", "mdsf_sync",
    "mdsf_read_sf", "mdsf_export_sf_device", "mdsf_export_plot_grids", "mdsf_cylindrical_average", "mdsf_reset", "mdsf_debug_cell_indices", "mdsf_debug_coords",
    "mdsf_debug_density", "mdsf_kernel_launches", "mdsf_frames_done", "mdsf_fft_path", "mdsf_splat_path",
    "mdsf_batch_frames", "mdsf_geometry", "mdsf_set_pretransform",
    "mdsf_enable_timing", "mdsf_stage_ms", "mdsf_timer_start", "mdsf_timer_stop", "mdsf_last_error",
    "mdsf_input_mark", "mdsf_input_wait",
    "mdsf_abi_version",
]


class MdsfError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libmdsf error %d: %s" % (code, message))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("n", C.c_int32 * 3), ("nborder", C.c_int32),
        ("dr", C.c_double * 3), ("box", C.c_double * 3), ("ucell", C.c_double * 9), ("ntypes", C.c_int32),
        ("amp", C.POINTER(C.c_double)), ("two_sig2", C.POINTER(C.c_double)), ("halfw", C.POINTER(C.c_int32)),
        ("coord_dtype", C.c_int32), ("arith_dtype", C.c_int32), ("fold_mode", C.c_int32), ("fft_mode", C.c_int32),
        ("batch_frames", C.c_int32), ("tile_x", C.c_int32), ("tile_y", C.c_int32), ("keep_density", C.c_int32),
        ("splat_mode", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


_lib = None


def load():
    """Load libmdsf.so once; raises ImportError with the build command if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with md-structure-factor_b200/csrc/build.sh "
                          "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
    sig = {
        "mdsf_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "mdsf_destroy": (C.c_int, [vp]),
        "mdsf_set_atoms": (C.c_int, [vp, i64, C.POINTER(i32)]),
        "mdsf_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
        "mdsf_host_free": (C.c_int, [vp]),
        "mdsf_host_register": (C.c_int, [vp, C.c_size_t]),
        "mdsf_host_unregister": (C.c_int, [vp]),
        "mdsf_push_frames": (C.c_int, [vp, vp, i64, dp, i64, i64, i32]),
        "mdsf_push_density": (C.c_int, [vp, dp, i64]),
        "mdsf_sync": (C.c_int, [vp]),
        "mdsf_read_sf": (C.c_int, [vp, dp]),
        "mdsf_export_sf_device": (C.c_int, [vp, vp]),
        "mdsf_export_plot_grids": (C.c_int, [vp, dp, dp, dp, dp, dp, dp, dp, dp, dp]),
        "mdsf_cylindrical_average": (C.c_int, [C.c_int, dp, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_int, dp, C.POINTER(i32), C.c_int, dp, dp]),
        "mdsf_reset": (C.c_int, [vp]),
        "mdsf_debug_cell_indices": (C.c_int, [vp, i64, C.POINTER(i32)]),
        "mdsf_debug_coords": (C.c_int, [vp, i64, dp]),
        "mdsf_debug_density": (C.c_int, [vp, i64, dp]),
        "mdsf_kernel_launches": (i64, [vp]),
        "mdsf_frames_done": (i64, [vp]),
        "mdsf_fft_path": (C.c_char_p, [vp]),
        "mdsf_splat_path": (C.c_char_p, [vp]),
        "mdsf_batch_frames": (C.c_int, [vp]),
        "mdsf_geometry": (C.c_int, [vp, C.POINTER(i32)]),
        "mdsf_set_pretransform": (C.c_int, [vp, i32, C.c_double, C.c_double]),
        "mdsf_enable_timing": (C.c_int, [vp, i32]),
        "mdsf_stage_ms": (C.c_int, [vp, dp, C.POINTER(i64)]),
        "mdsf_input_mark": (C.c_int, [vp, C.POINTER(C.c_int64)]),
        "mdsf_input_wait": (C.c_int, [vp, i64]),
        "mdsf_timer_start": (C.c_int, [vp]),
        "mdsf_timer_stop": (C.c_int, [vp, dp]),
        "mdsf_last_error": (C.c_char_p, []),
        "mdsf_abi_version": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.mdsf_abi_version() != ABI_VERSION:
        raise ImportError("libmdsf.so ABI %d != binding ABI %d" % (lib.mdsf_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise MdsfError(rc, load().mdsf_last_error().decode("utf-8", "replace"))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (cudaMallocHost), freed with the array."""
    lib = load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _check(lib.mdsf_host_alloc(max(nbytes, 1), C.byref(p)))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.mdsf_host_free, C.c_void_p(p.value))
    return arr


class Engine:
    """One GPU engine: per-frame path of dens.compute_sf (reference dens.py:277-321)."""

    def __init__(self, n, nborder, dr, box, ucell, amp, two_sig2, halfw, coord_dtype, arith_dtype,
                 fold_mode=FOLD_REFERENCE, fft_mode=FFT_AUTO, batch_frames=0, tile=(0, 0), keep_density=False,
                 device=0):
        self._lib = load()
        self._h = C.c_void_p()
        self.n = tuple(int(v) for v in n)
        self.coord_dtype = np.dtype(coord_dtype)
        if self.coord_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("coordinates must be float32 or float64")
        amp = np.ascontiguousarray(amp, dtype=np.float64)
        two_sig2 = np.ascontiguousarray(two_sig2, dtype=np.float64)
        halfw = np.ascontiguousarray(halfw, dtype=np.int32).reshape(-1, 3)
        cfg = Config()
        cfg.abi_version = ABI_VERSION
        cfg.device = int(device)
        self.device = int(device)
        cfg.n[:] = self.n
        cfg.nborder = int(nborder)
        cfg.dr[:] = [float(v) for v in dr]
        cfg.box[:] = [float(v) for v in box]
        cfg.ucell[:] = [float(v) for v in np.asarray(ucell, dtype=np.float64).reshape(9)]
        cfg.ntypes = int(amp.shape[0])
        cfg.amp = _dptr(amp)
        cfg.two_sig2 = _dptr(two_sig2)
        cfg.halfw = halfw.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.coord_dtype = F32 if self.coord_dtype == np.float32 else F64
        cfg.arith_dtype = F32 if np.dtype(arith_dtype) == np.float32 else F64
        cfg.fold_mode = int(fold_mode)
        cfg.fft_mode = int(fft_mode)
        cfg.batch_frames = int(batch_frames)
        cfg.tile_x, cfg.tile_y = int(tile[0]), int(tile[1])
        cfg.keep_density = 1 if keep_density else 0
        cfg.splat_mode = 0
        _check(self._lib.mdsf_create(C.byref(cfg), C.byref(self._h)))
        self.natoms = 0
        self._finalizer = weakref.finalize(self, self._lib.mdsf_destroy, C.c_void_p(self._h.value))

    # -- lifetime
    def close(self):
        if self._h is not None and self._finalizer.alive:
            self._finalizer()
        self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- data
    def set_atoms(self, type_ids):
        t = np.ascontiguousarray(type_ids, dtype=np.int32)
        _check(self._lib.mdsf_set_atoms(self._h, t.shape[0], t.ctypes.data_as(C.POINTER(C.c_int32))))
        self.natoms = int(t.shape[0])

    def push_frames(self, coords, scale, wrap_range=None, write_back=False):
        """coords: C-contiguous (T, Na, 3) array of the engine's coord dtype; scale: (T, 3)."""
        if coords.dtype != self.coord_dtype or not coords.flags.c_contiguous:
            raise TypeError("coords must be C-contiguous %s" % self.coord_dtype)
        if coords.ndim != 3 or coords.shape[1] != self.natoms or coords.shape[2] != 3:
            raise ValueError("coords must have shape (T, %d, 3), got %s" % (self.natoms, coords.shape))
        if write_back and not coords.flags.writeable:
            raise ValueError("write_back needs a writeable coords array")
        scale = np.ascontiguousarray(scale, dtype=np.float64).reshape(coords.shape[0], 3)
        lo, hi = (0, self.natoms) if wrap_range is None else wrap_range
        _check(self._lib.mdsf_push_frames(self._h, C.c_void_p(coords.ctypes.data), coords.shape[0], _dptr(scale),
                                          int(lo), int(hi), 1 if write_back else 0))

    def push_frames_ptr(self, ptr, nframes, scale, wrap_range=None, write_back=False):
        """Same as push_frames for a raw host or DEVICE pointer to (nframes, Na, 3) coordinates."""
        scale = np.ascontiguousarray(scale, dtype=np.float64).reshape(nframes, 3)
        lo, hi = (0, self.natoms) if wrap_range is None else wrap_range
        _check(self._lib.mdsf_push_frames(self._h, C.c_void_p(int(ptr)), int(nframes), _dptr(scale), int(lo), int(hi),
                                          1 if write_back else 0))

    def push_density(self, d1):
        d1 = np.ascontiguousarray(d1, dtype=np.float64)
        if d1.ndim == 3:
            d1 = d1[None]
        if tuple(d1.shape[1:]) != self.n:
            raise ValueError("density must have shape (T,) + %s" % (self.n,))
        _check(self._lib.mdsf_push_density(self._h, _dptr(d1), d1.shape[0]))

    def sync(self):
        _check(self._lib.mdsf_sync(self._h))

    def mark_input(self):
        """Ticket behind every host<->device copy queued so far (see wait_input)."""
        t = C.c_int64()
        _check(self._lib.mdsf_input_mark(self._h, C.byref(t)))
        return t.value

    def wait_input(self, ticket):
        """Block until the copies behind ``ticket`` are done: the host buffers pushed before it can be reused."""
        _check(self._lib.mdsf_input_wait(self._h, int(ticket)))

    def read_sf(self, out=None):
        """sf (Nx, Ny, Nz/2+1) float64, into ``out`` if given (e.g. a pinned_empty buffer for a faster copy)."""
        shape = (self.n[0], self.n[1], self.n[2] // 2 + 1)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape %s" % (shape,))
        _check(self._lib.mdsf_read_sf(self._h, _dptr(out)))
        return out

    def export_plot_grids(self, kaxes, paxes, want=("sfplt", "kgrid", "kgridplt")):
        """sfplt / kgrid / kgridplt of reference dens.py:323-344, assembled on the GPU from the resident S(q).
        ``kaxes`` = (kx[Nx], ky[Ny], kz[M]), ``paxes`` = (px[Nx-2], py[Ny-2], pz[2M-3]): per-index axis values."""
        nx, ny, m = self.n[0], self.n[1], self.n[2] // 2 + 1
        ax = [np.ascontiguousarray(a, dtype=np.float64) for a in tuple(kaxes) + tuple(paxes)]
        if [a.size for a in ax] != [nx, ny, m, nx - 2, ny - 2, 2 * m - 3]:
            raise ValueError("axis vectors do not match the grid")
        out = {"sfplt": np.empty((nx - 2, ny - 2, 2 * m - 3)) if "sfplt" in want else None,
               "kgrid": np.empty((nx, ny, m, 4)) if "kgrid" in want else None,
               "kgridplt": np.empty((nx - 2, ny - 2, 2 * m - 3, 4)) if "kgridplt" in want else None}
        ptr = lambda a: _dptr(a) if a is not None else None
        _check(self._lib.mdsf_export_plot_grids(self._h, *[_dptr(a) for a in ax], ptr(out["sfplt"]), ptr(out["kgrid"]), ptr(out["kgridplt"])))
        return out

    def export_sf_device(self, device_ptr):
        _check(self._lib.mdsf_export_sf_device(self._h, C.c_void_p(int(device_ptr))))

    def reset(self):
        _check(self._lib.mdsf_reset(self._h))

    # -- parity taps
    def debug_cell_indices(self, frame):
        out = np.empty((self.natoms, 3), dtype=np.int32)
        _check(self._lib.mdsf_debug_cell_indices(self._h, int(frame), out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def debug_coords(self, frame):
        out = np.empty((self.natoms, 3), dtype=np.float64)
        _check(self._lib.mdsf_debug_coords(self._h, int(frame), _dptr(out)))
        return out

    def debug_density(self, frame):
        out = np.empty(self.n, dtype=np.float64)
        _check(self._lib.mdsf_debug_density(self._h, int(frame), _dptr(out)))
        return out

    # -- introspection
    @property
    def kernel_launches(self):
        return int(self._lib.mdsf_kernel_launches(self._h))

    @property
    def frames_done(self):
        return int(self._lib.mdsf_frames_done(self._h))

    @property
    def fft_path(self):
        return self._lib.mdsf_fft_path(self._h).decode()

    @property
    def splat_path(self):
        return self._lib.mdsf_splat_path(self._h).decode()

    @property
    def batch_frames(self):
        return int(self._lib.mdsf_batch_frames(self._h))

    def set_pretransform(self, theta):
        """Apply the CLI's monoclinic transform (reference main_gromacs.py:204-207) inside the first kernel to every frame
        pushed from now on; ``theta`` in radians, ``None`` switches it off.  np.sin / np.cos give the float64 scalars
        the reference divides / multiplies by."""
        if theta is None:
            _check(self._lib.mdsf_set_pretransform(self._h, 0, 1.0, 0.0))
        else:
            _check(self._lib.mdsf_set_pretransform(self._h, 1, float(np.sin(theta)), float(np.cos(theta))))

    @property
    def geometry(self):
        """Launch geometry: splat tile, z-slab width, volume layout chunk width, pass tile widths, splat smem (KiB)."""
        g = (C.c_int32 * 8)()
        _check(self._lib.mdsf_geometry(self._h, g))
        return {"tile": (int(g[0]), int(g[1])), "slab": int(g[2]), "nslab": int(g[3]), "layout_w": int(g[4]),
                "wy": int(g[5]), "wx": int(g[6]), "splat_smem_kib": int(g[7])}

    def enable_timing(self, on=True):
        _check(self._lib.mdsf_enable_timing(self._h, 1 if on else 0))

    def timer_start(self):
        _check(self._lib.mdsf_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        _check(self._lib.mdsf_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def stage_ms(self):
        out = np.zeros(6)
        nb = C.c_int64()
        _check(self._lib.mdsf_stage_ms(self._h, _dptr(out), C.byref(nb)))
        names = ["copy", "prep_bin", "splat_zfft", "fft_y", "fft_x_accum", "total"]
        return dict(zip(names, out.tolist())), int(nb.value)
