"""Drop-in replacement for the reference's ``dens.py`` with the frame loop on a B200.

Same public surface as joeyelk/MD-Structure-Factor ``dens.py``:
``compute_sf(r, L, typ, out_filename, rad, ucell, Sres)`` (reference dens.py:166-346),
``load_radii`` (:23-35), ``get_borders`` (:38-43), ``rescale`` (:46-62), ``remap_grid_tcl``
(:86-108), ``get_dplot`` (:142-163) and the module knobs ``USE_BETTER_RESOLUTION``,
``PRINT_DETAILS``, ``RANDOM_NOISE``, ``Nspatialgrid``, ``theta``, ``PRECISION``, ``dtyp`` (:10-20).

What differs: the per-frame work (rescale, PBC wrap, cell index, Gaussian stamp, periodic fold,
3-D FFT, |F|^2 accumulation; reference dens.py:277-321) runs in libmdsf.so on the GPU.  Everything
that is computed once per call (grid sizing, half widths, plot grids, the npz file) is done
here with the reference's numpy expressions, so those values are identical by construction.
There is no CPU implementation of the frame loop in this package.
"""
import math

import numpy as np

import mdsf_native as _native
import npz_writer

USE_BETTER_RESOLUTION = True
PRINT_DETAILS = True
RANDOM_NOISE = 0          # > 0: replace the density by np.random.rand noise (reference dens.py:268-280)

Nspatialgrid = np.asarray([0.0, 0.0, 0.0])
theta = np.pi / 3

PRECISION = 1.0E-24       # amplitude at which a Gaussian stamp is cut off

dtyp = np.float64         # kept for API compatibility; the engine always works in float64

# ---- knobs that only exist in this implementation
FOLD_MODE = "reference"   # "reference": reproduce the corner rule of reference dens.py:107; "periodic": exact fold
WRITE_BACK_COORDS = True  # the reference rescales/wraps ``r`` in place; keep that side effect
DEVICE = 0                # CUDA device ordinal used by compute_sf
FFT_MODE = "auto"         # "auto" | "native" | "cufft"
BATCH_FRAMES = 0          # frames per device batch (0 = automatic)
SAVE_COMPRESSED = True    # np.savez_compressed like the reference; False writes an uncompressed npz
GPU_MONOCLINIC = True     # compute_sf_stream: the monoclinic transform of main_gromacs.py:204-207 runs inside the first kernel, not in numpy
GPU_PLOT_GRIDS = True     # sfplt / kgrid / kgridplt (dens.py:323-344) assembled on the GPU from the resident S(q); False = numpy
PARALLEL_NPZ = True       # deflate the npz members on all host cores (same container, np.load reads it); False = numpy's writer
LAST_RUN = {}             # filled by compute_sf: grid, batch size, FFT path, kernel launches

_BUFFSIZE = 1000000       # reference dens.py:206


def load_radii(filename):
    """Read ``Z  label  radius_pm`` rows -> {label: (electrons, sigma in Angstrom)} (reference dens.py:23-35)."""
    table = {}
    with open(filename) as handle:
        for row in handle:
            fields = row.strip().split()
            table[fields[1]] = (float(fields[0]), float(fields[2]) / 100.0)
    return table


def get_borders(ad, dr, keys):
    """Half width of each label's Gaussian stamp in grid steps (reference dens.py:38-43)."""
    return {key: ad[key][1] * np.sqrt(np.log(ad[key][0] / PRECISION)) / dr for key in keys}


def rescale(coords, dims):
    """Scale every frame to the mean box, in place (reference dens.py:46-62); host-side helper.

    compute_sf() does not call this: the engine applies the same multiplication on the GPU."""
    avgdims = np.average(dims, axis=0)
    factor = avgdims / dims
    for it in range(coords.shape[0]):
        for i in range(3):
            coords[it, :, i] *= factor[it, i]
    return coords, avgdims


def remap_grid_tcl(d0, des, ori):
    """Fold the padding of ``d0`` back onto the periodic cell (reference dens.py:86-108).

    ``ori[d]`` / ``des[d]`` are the four slice bounds per dimension of the padded array and of the
    cell.  Corner regions use the y-side block for their z destination, as the reference does
    (dens.py:107).  Host-side helper; the engine folds on the fly inside the splat kernel."""
    d1 = np.copy(d0[ori[0][1]:ori[0][2], ori[1][1]:ori[1][2], ori[2][1]:ori[2][2]])

    def src(dim, block):
        return slice(ori[dim][block], ori[dim][block + 1])

    def dst(dim, block):
        return slice(des[dim][2 - block], des[dim][3 - block]) if block != 1 else slice(None)

    for bx in (0, 1, 2):
        for by in (0, 1, 2):
            for bz in (0, 1, 2):
                if bx == by == bz == 1:
                    continue
                zdst = dst(2, by) if (bx != 1 and by != 1 and bz != 1) else dst(2, bz)
                d1[dst(0, bx), dst(1, by), zdst] += d0[src(0, bx), src(1, by), src(2, bz)]
    return d1


def get_dplot(dmag):
    """Half spectrum -> centred, cropped full-spectrum view (reference dens.py:142-163)."""
    nx, ny, m = dmag.shape
    hz = m - 1
    full = np.zeros((nx, ny, 2 * m - 1))
    full[:, :, hz:2 * hz] = np.roll(dmag[:, :, :hz], (nx // 2, ny // 2), axis=(0, 1))
    full[1:nx - 1, 1:ny - 1, :hz] = full[:1:-1, :1:-1, 2 * hz:hz:-1]
    cx, cy, cz = full.shape[0] // 2, full.shape[1] // 2, full.shape[2] // 2
    full[cx, cy, cz] = 1.0 / 3.0 * (full[cx + 1, cy, cz] + full[cx, cy + 1, cz] + full[cx, cy, cz + 1])
    return full[1:-1, 1:-1, 1:-1]


def _wrapped_atoms(nframes, natoms):
    """Atom index range the reference's PBC pass touches (dens.py:209-221).

    Systems with >= 1e6 atoms take the buffered branch, which only wraps r[it, imin:nframes]."""
    if natoms < _BUFFSIZE:
        return 0, natoms
    lo = hi = 0
    for imin in range(0, nframes, _BUFFSIZE):
        if imin + _BUFFSIZE > nframes:
            lo, hi = imin, min(nframes, natoms)
    return lo, hi


def _grid(L, Sres):
    """Grid size and spacing (reference dens.py:181-189, 202)."""
    n = (L / Sres).astype(int)
    if USE_BETTER_RESOLUTION:
        for i in range(3):
            if L[i] / n[i] > Sres:
                n[i] += 1
    for i in range(3):
        n[i] += n[i] % 2
    return n, np.divide(L, n)


def _k_axes(shape, L):
    """Per-index axis values of kgrid (reference dens.py:327-334) and kgridplt (dens.py:337-342): the reference's scalar
    expressions, one evaluation per index.  Indices kgrid never writes (the middle one of an odd axis) stay 0."""
    nx, ny, m = shape
    vx, vy, vz = np.zeros(nx), np.zeros(ny), np.zeros(m)
    for ix in range(int(nx / 2)):
        vx[ix] = ix * 2.0 * math.pi / L[0]
        vx[nx - 1 - ix] = -(ix + 0.5) * 2.0 * math.pi / L[0]
    for iy in range(int(ny / 2)):
        vy[iy] = iy * 2.0 * math.pi / L[1]
        vy[ny - 1 - iy] = -(iy + 0.5) * 2.0 * math.pi / L[1]
    for iz in range(m):
        vz[iz] = iz * 2.0 * math.pi / L[2]
    paxes = []
    for d, nd in enumerate((nx - 2, ny - 2, m * 2 - 3)):
        paxes.append(np.asarray([(i - nd / 2) * 2.0 * math.pi / L[d] for i in range(nd)], dtype=np.float64))
    return (vx, vy, vz), tuple(paxes)


def _k_lattices(shape, L):
    """kgrid / kgridplt coordinate lattices (reference dens.py:325-342), host route: one broadcast per channel."""
    nx, ny, m = shape
    (vx, vy, vz), paxes = _k_axes(shape, L)
    kgrid = np.zeros((nx, ny, m, 4))
    kgrid[..., 0] = vx[:, None, None]
    kgrid[..., 1] = vy[None, :, None]
    kgrid[..., 2] = vz[None, None, :]
    kplt = np.zeros((nx - 2, ny - 2, m * 2 - 3, 4))
    for d in range(3):
        view = [None, None, None]
        view[d] = slice(None)
        kplt[..., d] = paxes[d][tuple(view)]
    return kgrid, kplt


def make_engine(L_mean, typ, rad, ucell, Sres, coord_dtype, arith_dtype, keep_density=False, device=None,
                batch_frames=None, fft_mode=None, tile=(0, 0), fold_mode=None):
    """Build the GPU engine for one call: everything that is frame-invariant (dens.py:181-231).

    Returns (engine, N, dr).  Raises KeyError for labels missing from ``rad`` like the reference."""
    n, dr = _grid(L_mean, Sres)
    typ = np.asarray(typ)
    labels, type_ids = np.unique(typ.astype(str), return_inverse=True)
    bdict = get_borders(rad, dr, set(labels.tolist()))        # KeyError for unknown labels (dens.py:42)
    nborder = int(max(max(v) for v in bdict.values()))          # dens.py:226-231
    amp = np.array([rad[l][0] / np.power(rad[l][1], 3.) for l in labels])       # dens.py:308
    two_sig2 = np.array([2 * np.power(rad[l][1], 2.) for l in labels])
    halfw = np.array([bdict[l].astype(int) for l in labels])                    # dens.py:287
    fold = {"reference": _native.FOLD_REFERENCE, "periodic": _native.FOLD_PERIODIC}[fold_mode or FOLD_MODE]
    fft = {"auto": _native.FFT_AUTO, "native": _native.FFT_NATIVE, "cufft": _native.FFT_CUFFT}[fft_mode or FFT_MODE]
    eng = _native.Engine(n, nborder, dr, L_mean, ucell, amp, two_sig2, halfw, coord_dtype, arith_dtype,
                         fold_mode=fold, fft_mode=fft, batch_frames=BATCH_FRAMES if batch_frames is None else batch_frames,
                         tile=tile, keep_density=keep_density, device=DEVICE if device is None else device)
    eng.set_atoms(type_ids)
    return eng, n, dr, nborder


def finish_sf(sf, L, n, out_filename, eng=None):
    """Plot grids and the sf npz (reference dens.py:323-346).  With a live engine (and GPU_PLOT_GRIDS) sfplt, kgrid and
    kgridplt are assembled on the GPU from the resident S(q) (mdsf_export_plot_grids: gather + broadcast, bit-identical);
    otherwise -- e.g. after the NCCL reduce of sf_distributed -- by the numpy expressions."""
    if eng is not None and GPU_PLOT_GRIDS and min(sf.shape[0], sf.shape[1]) >= 4 and sf.shape[2] >= 3:
        kaxes, paxes = _k_axes(sf.shape, L)
        g = eng.export_plot_grids(kaxes, paxes)
        sfplt, kgrid, kgridplt = g["sfplt"], g["kgrid"], g["kgridplt"]
    else:
        sfplt = get_dplot(sf)
        kgrid, kgridplt = _k_lattices(sf.shape, L)
        kgridplt[:, :, :, 3] = sfplt
    if PARALLEL_NPZ:      # same zip-of-npy container as np.savez_compressed, deflated on all cores (npz_writer.py)
        npz_writer.savez_parallel(out_filename, compressed=SAVE_COMPRESSED, sf=sf, sfplt=sfplt, L=L, N=n, kgrid=kgrid, kgridplt=kgridplt)
    else:
        save = np.savez_compressed if SAVE_COMPRESSED else np.savez
        save(out_filename, sf=sf, sfplt=sfplt, L=L, N=n, kgrid=kgrid, kgridplt=kgridplt)


def compute_sf_stream(frames, L, typ, out_filename, rad, ucell, Sres, first_frame=0, end_frame=None, monoclinic_theta=None,
                      chunk_frames=None):
    """``compute_sf`` for a frame SOURCE instead of an in-memory array (north_star: the loader streams frames into
    pinned, double-buffered host memory overlapped with the H2D copies).

    ``frames``: object with ``shape`` (T, Na, 3), ``dtype``, ``read_into(buf) -> k`` and ``skip(n)``, e.g.
    ``load_traj.NpzFrameStream``.  ``L``: all box lengths (T, 3).  Frames ``[first_frame:end_frame]`` are used, like the
    slice main_gromacs.py:211 passes.  ``monoclinic_theta`` (radians) applies the reference's coordinate transform
    (main_gromacs.py:204-207) chunk by chunk -- the same two numpy expressions, so the same bits.
    Two pinned chunk buffers alternate: while the GPU works on one, the next chunk is decoded into the other; a
    buffer is reused as soon as its host-to-device copies are done (mdsf_input_mark / mdsf_input_wait)."""
    global Nspatialgrid
    T, natoms = int(frames.shape[0]), int(frames.shape[1])
    lo_f, hi_f, _ = slice(first_frame, end_frame).indices(T)
    dims = np.asarray(L)[lo_f:hi_f]
    if dims.dtype not in (np.float32, np.float64):
        dims = dims.astype(np.float64)
    cdtype = np.dtype(frames.dtype)
    if cdtype not in (np.float32, np.float64):
        raise TypeError("streamed coordinates must be float32 or float64")
    arith = np.float32 if (cdtype == np.float32 and dims.dtype == np.float32) else np.float64
    Lm = np.average(dims, axis=0)
    scale = (Lm / dims).astype(np.float64)
    nframes = hi_f - lo_f
    eng, n, dr, nborder = make_engine(Lm, typ, rad, ucell, Sres, cdtype, arith)
    Nspatialgrid = n
    try:
        print("Calculating Structure factor for ", natoms, " atoms over ", nframes, " timesteps (streamed). \n", "Progress: ")
        print("GPU engine: batch of %d frames, %s splat, %s FFT, border %d cells" % (eng.batch_frames, eng.splat_path, eng.fft_path, nborder))
        chunk = int(chunk_frames or 2 * eng.batch_frames)
        if monoclinic_theta is not None and GPU_MONOCLINIC:
            eng.set_pretransform(monoclinic_theta)      # K1 does main_gromacs.py:206-207 (same float64 expressions, same bits)
        bufs = [_native.pinned_empty((chunk, natoms, 3), cdtype) for _ in range(2)]
        tickets = [None, None]
        wrap = _wrapped_atoms(nframes, natoms)
        frames.skip(lo_f)
        done, i = 0, 0
        while done < nframes:
            b = i % 2
            if tickets[b] is not None:
                eng.wait_input(tickets[b])          # the copies out of this buffer are done; the kernels may still run
            k = frames.read_into(bufs[b][:min(chunk, nframes - done)])
            if k == 0:
                raise EOFError("frame source ended after %d of %d frames" % (done, nframes))
            blk = bufs[b][:k]
            if monoclinic_theta is not None and not GPU_MONOCLINIC:
                blk[..., 1] = blk[..., 1] / np.sin(monoclinic_theta)
                blk[..., 0] = blk[..., 0] - blk[..., 1] * np.cos(monoclinic_theta)
            eng.push_frames(blk, scale[done:done + k], wrap)
            tickets[b] = eng.mark_input()
            done += k
            i += 1
        sf = eng.read_sf()
        LAST_RUN.clear()
        LAST_RUN.update(N=n.copy(), dr=dr.copy(), Nborder=nborder, batch_frames=eng.batch_frames, fft=eng.fft_path, splat=eng.splat_path,
                        kernel_launches=eng.kernel_launches, frames=eng.frames_done, streamed_chunks=i, chunk_frames=chunk)
        finish_sf(sf, Lm, n, out_filename, eng)
    finally:
        eng.close()


def compute_sf(r, L, typ, out_filename, rad, ucell, Sres):
    """
    compute 3d structure factor (same contract as the reference, dens.py:166-178)
    :param r: coordinates (T, Na, 3); rescaled and wrapped in place like the reference does
    :param L: cell dimensions (T, 3)
    :param typ: list of element types or names, keys into ``rad``
    :param out_filename: output filename; ``out_filename + '.npz'`` receives sf, sfplt, L, N, kgrid, kgridplt
    :param rad: dictionary from load_radii
    :param ucell: unit cell as a 3x3 numpy array
    :param Sres: spatial resolution of the density grid in angstroms
    """
    global Nspatialgrid
    r_in = r
    dims = np.asarray(L)
    if dims.dtype not in (np.float32, np.float64):
        dims = dims.astype(np.float64)
    if r.dtype not in (np.float32, np.float64) or not r.flags.c_contiguous:
        r = np.ascontiguousarray(r, dtype=r.dtype if r.dtype in (np.float32, np.float64) else np.float64)
    arith = np.float32 if (r.dtype == np.float32 and dims.dtype == np.float32) else np.float64

    L = np.average(dims, axis=0)          # dens.py:52
    scale = L / dims                      # dens.py:53; applied per frame by the engine
    nframes, natoms = r.shape[0], r.shape[1]

    eng, n, dr, nborder = make_engine(L, typ, rad, ucell, Sres, r.dtype, arith)
    Nspatialgrid = n
    try:
        if PRINT_DETAILS:
            print("=" * 50)
            print("=" * 50)
            print("Unit cell has dimensions of %s" % L)
            print("requested resolution: %s Angstroms" % Sres)
            print("Using a spatial grid of: %s" % n)
            print("actual resolution: %s Angstroms" % (L / n))
            print("This corresponds to maximum q vector of %s inverse Angstroms" % (2 * math.pi * n / L))
            print("=" * 50)
            print("=" * 50)
        print("remapping coordinates into periodic cell")
        print("allocating memory")
        print(dtyp)
        print((int(n[0]), int(n[1]), int(n[2]) / 2 + 1))
        print("Calculating Structure factor for ", natoms, " atoms over ", nframes, " timesteps. \n", "Progress: ")
        print("GPU engine: batch of %d frames, %s splat, %s FFT, border %d cells" % (eng.batch_frames, eng.splat_path, eng.fft_path, nborder))

        if RANDOM_NOISE > 0:
            print("*" * 80)
            print("*" * 80)
            print("OVERRIDING DENSITY WITH RANDOM NOISE")
            print(nframes * RANDOM_NOISE, "timesteps")
            print("*" * 80)
            print("*" * 80)
            for it in range(nframes):     # the reference draws one noise volume per loop iteration (dens.py:277-280)
                eng.push_density(np.random.rand(int(n[0]), int(n[1]), int(n[2])))
        else:
            lo, hi = _wrapped_atoms(nframes, natoms)
            eng.push_frames(r, scale.astype(np.float64), (lo, hi), write_back=WRITE_BACK_COORDS)
        sf = eng.read_sf()
        LAST_RUN.clear()
        LAST_RUN.update(N=n.copy(), dr=dr.copy(), Nborder=nborder, batch_frames=eng.batch_frames, fft=eng.fft_path, splat=eng.splat_path,
                        kernel_launches=eng.kernel_launches, frames=eng.frames_done)
        if WRITE_BACK_COORDS and r is not r_in and RANDOM_NOISE <= 0:
            r_in[...] = r
        finish_sf(sf, L, n, out_filename, eng)
    finally:
        eng.close()
