"""CPU oracle for the structure-factor hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy *restatement* of the algorithm in the reference's
``dens.py`` (joeyelk/MD-Structure-Factor).  It exists so that the CUDA path can
be checked on machines where ``/root/reference`` is not present (the GPU box).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it; the product package never does.

Parity pin: the reference ships no golden vectors for this path (its only test
is ``test/test_load_traj.py``), so the oracle is pinned against outputs of the
*unmodified reference run in the build container*: ``tests/golden/make_golden.py``
imports ``/root/reference/dens.py`` under a shim for ``past.utils.old_div`` and
commits inputs + outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against them.

Every function cites the reference lines it follows (paths are relative to the
reference checkout).
"""
import math

import numpy as np

PRECISION = 1.0e-24      # dens.py:18
WRAP_BUFFSIZE = 1000000  # dens.py:206


# --------------------------------------------------------------------------- tables
def load_radii(path):
    """dens.py:23-35 -- ``Z  label  radius_pm`` rows -> {label: (Nel, sigma_A)}; later rows win."""
    table = {}
    with open(path) as fh:
        for raw in fh:
            cols = raw.strip().split()
            table[cols[1]] = (float(cols[0]), float(cols[2]) / 100.0)
    return table


def half_widths(rad, dr, labels, precision=PRECISION):
    """dens.py:38-43 -- real-valued stamp half width (in cells, per dimension) for each label."""
    out = {}
    for lab in labels:
        nel, sigma = rad[lab]
        out[lab] = sigma * np.sqrt(np.log(nel / precision)) / dr
    return out


def border_cells(widths):
    """dens.py:226-231 -- padding thickness B = int(max over labels and dims)."""
    return int(max(max(w) for w in widths.values()))


# --------------------------------------------------------------------------- geometry
def rescale_frames(coords, dims):
    """dens.py:46-62 -- scale every frame to the mean box, IN PLACE, in the coords dtype.

    Returns the mean box (dtype of ``dims``)."""
    mean_box = np.average(dims, axis=0)
    factor = mean_box / dims
    for t in range(coords.shape[0]):
        for d in range(3):
            coords[t, :, d] *= factor[t, d]
    return mean_box


def grid_shape(box, sres, better_resolution=True):
    """dens.py:181-189,202 -- even grid with spacing <= sres; returns (N int64[3], dr float64[3])."""
    n = (box / sres).astype(int)
    if better_resolution:
        for d in range(3):
            if box[d] / n[d] > sres:
                n[d] += 1
    for d in range(3):
        n[d] += n[d] % 2
    return n, np.divide(box, n)


def wrapped_atom_range(nframes, natoms, buffsize=WRAP_BUFFSIZE):
    """dens.py:209-221 -- which atom indices the reference's PBC pass touches.

    Below ``buffsize`` atoms every atom is wrapped.  At or above it the reference walks
    ``range(0, nframes, buffsize)`` and wraps ``[imin, nframes)`` only inside the
    ``imax > nframes`` branch, i.e. only the first ``nframes`` atoms (dens.py:216-221)."""
    if natoms < buffsize:
        return 0, natoms
    lo, hi = 0, 0
    for imin in range(0, nframes, buffsize):
        if imin + buffsize > nframes:
            lo, hi = imin, min(nframes, natoms)
    return lo, hi


def wrap_frames(coords, box):
    """dens.py:209-221 -- single +-L shift, in place: r>=L -> r-L, then r<=0 -> r+L."""
    lo, hi = wrapped_atom_range(coords.shape[0], coords.shape[1])
    zero = [0.0, 0.0, 0.0]
    for t in range(coords.shape[0]):
        blk = coords[t, lo:hi, :]
        blk = np.where(blk < box, blk, blk - box)
        coords[t, lo:hi, :] = blk
        blk = coords[t, lo:hi, :]
        coords[t, lo:hi, :] = np.where(blk > zero, blk, blk + box)


def cell_indices(frame, dr):
    """dens.py:285 -- trunc(r/dr) with an fp64 divide, for one frame (Na,3) -> int64."""
    return (frame / dr).astype(int)


# --------------------------------------------------------------------------- density
def stamp_padded(frame, labels, rad, widths, n, dr, nb, ucell, out=None):
    """dens.py:283-308 -- truncated Gaussian of every atom on the zero-padded grid.

    Stamp box per dimension is [ir-A+B, ir+A+B) (2A cells, asymmetric about the floor cell,
    dens.py:292-297); displacement is r - (i-B)*dr (dens.py:252-256,299); Cartesian
    components are sum_m ucell[m,l]*b_m (einsum 'ml,ijkm', dens.py:301); amplitude is
    Nel/sigma^3 * exp(-|c|^2/(2 sigma^2)) (dens.py:303-308)."""
    shape = tuple(int(n[d]) + 2 * nb for d in range(3))
    d0 = np.zeros(shape) if out is None else out
    u = np.asarray(ucell, dtype=np.float64)
    for a in range(frame.shape[0]):
        r = frame[a, :]
        ir = (r / dr).astype(int)
        hw = widths[labels[a]].astype(int)
        lo = ir - hw + nb
        hi = ir + hw + nb
        b = [r[d] - (np.arange(lo[d], hi[d]) - nb) * dr[d] for d in range(3)]
        bx = b[0][:, None, None]
        by = b[1][None, :, None]
        bz = b[2][None, None, :]
        dist2 = None
        for l in range(3):
            c = u[0, l] * bx + u[1, l] * by + u[2, l] * bz
            dist2 = c * c if dist2 is None else dist2 + c * c
        nel, sigma = rad[labels[a]]
        d0[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += (nel / np.power(sigma, 3.0)) * np.exp(
            -dist2 / (2 * np.power(sigma, 2.0)))
    return d0


def fold_padded(d0, n, nb, mode="reference"):
    """dens.py:86-108 -- add the padding back onto the periodic cell.

    ``mode='reference'`` reproduces the corner rule of dens.py:107, where the destination
    z block of the 8 corner regions is selected by the *y* side: a corner cell whose y side
    differs from its z side lands at z = p_z + B (z low) or p_z - B (z high) instead of
    p_z +- N_z.  ``mode='periodic'`` is the mathematically periodic fold."""
    nx, ny, nz = (int(v) for v in n)
    out = d0[nb:nb + nx, nb:nb + ny, nb:nb + nz].copy()

    def seg(side, nn):
        # (source slice in padded array, destination slice in the cell) for one dimension
        if side == 0:
            return slice(nb, nb + nn), slice(0, nn)
        if side < 0:
            return slice(0, nb), slice(nn - nb, nn)
        return slice(nb + nn, 2 * nb + nn), slice(0, nb)

    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sz in (-1, 0, 1):
                if sx == 0 and sy == 0 and sz == 0:
                    continue
                srcx, dstx = seg(sx, nx)
                srcy, dsty = seg(sy, ny)
                srcz, dstz = seg(sz, nz)
                if mode == "reference" and sx != 0 and sy != 0 and sz != 0:
                    dstz = seg(sy, nz)[1]   # dens.py:107 uses id2 for the z destination
                out[dstx, dsty, dstz] += d0[srcx, srcy, srcz]
    return out


def density_frame(frame, labels, rad, widths, n, dr, nb, ucell, fold_mode="reference"):
    """dens.py:283-311 for one (already rescaled and wrapped) frame -> periodic density d1."""
    d0 = stamp_padded(frame, labels, rad, widths, n, dr, nb, ucell)
    return fold_padded(d0, n, nb, fold_mode)


# --------------------------------------------------------------------------- spectrum
def power_spectrum(d1):
    """dens.py:313-316 -- |rfftn(d1)|^2."""
    f = np.fft.rfftn(d1)
    return f.real * f.real + f.imag * f.imag


def centred_view(sf):
    """dens.py:142-163 -- half spectrum -> centred, cropped full view (Nx-2, Ny-2, Nz-1).

    kz >= 0 half: quadrant swap in x,y (Nyquist kz plane dropped, dens.py:155-158);
    kz < 0 half: point inversion of the kz > 0 half (dens.py:160); the DC bin is replaced by
    the mean of its +x,+y,+z neighbours (dens.py:161); one cell is cropped from every face
    (dens.py:163)."""
    nx, ny, m = sf.shape
    nzh = m - 1
    full = np.zeros((nx, ny, 2 * m - 1))
    upper = np.roll(np.roll(sf[:, :, :nzh], nx // 2, axis=0), ny // 2, axis=1)
    full[:, :, nzh:2 * nzh] = upper
    # inversion: (X,Y,Z) <- (Nx-X, Ny-Y, Nz-Z) for X in [1,Nx-1), Y in [1,Ny-1), Z in [0,Nz/2)
    src = full[:1:-1, :1:-1, 2 * nzh:nzh:-1]
    full[1:nx - 1, 1:ny - 1, :nzh] = src
    cx, cy, cz = nx // 2, ny // 2, (2 * m - 1) // 2
    full[cx, cy, cz] = 1.0 / 3.0 * (full[cx + 1, cy, cz] + full[cx, cy + 1, cz] + full[cx, cy, cz + 1])
    return full[1:-1, 1:-1, 1:-1]


def k_lattices(sf_shape, box):
    """dens.py:325-342 -- kgrid (half spectrum) and kgridplt (centred view) coordinates.

    kgrid uses +i*2pi/L for i < N/2 and -(i+0.5)*2pi/L at the mirrored index (dens.py:327-334);
    channel 3 of kgrid stays zero.  Divisions are by the numpy scalar box[d], so the values
    carry the precision of the box dtype exactly as in the reference."""
    nx, ny, m = sf_shape
    kgrid = np.zeros((nx, ny, m, 4))
    for i in range(nx // 2):
        kgrid[i, :, :, 0] = i * 2.0 * math.pi / box[0]
        kgrid[nx - 1 - i, :, :, 0] = -(i + 0.5) * 2.0 * math.pi / box[0]
    for i in range(ny // 2):
        kgrid[:, i, :, 1] = i * 2.0 * math.pi / box[1]
        kgrid[:, ny - 1 - i, :, 1] = -(i + 0.5) * 2.0 * math.pi / box[1]
    for i in range(m):
        kgrid[:, :, i, 2] = i * 2.0 * math.pi / box[2]
    px, py, pz = nx - 2, ny - 2, 2 * m - 3
    kplt = np.zeros((px, py, pz, 4))
    for i in range(px):
        kplt[i, :, :, 0] = (i - px / 2) * 2.0 * math.pi / box[0]
    for i in range(py):
        kplt[:, i, :, 1] = (i - py / 2) * 2.0 * math.pi / box[1]
    for i in range(pz):
        kplt[:, :, i, 2] = (i - pz / 2) * 2.0 * math.pi / box[2]
    return kgrid, kplt


# --------------------------------------------------------------------------- whole path
def structure_factor(coords, dims, labels, rad, ucell, sres, fold_mode="reference",
                     better_resolution=True, taps=None, frame_callback=None):
    """dens.py:166-346 without the file write: returns the six arrays of the sf npz.

    ``coords`` (T,Na,3) is rescaled and wrapped in place like the reference does.
    ``taps`` (optional dict) receives per-stage intermediates: 'ir' (T,Na,3), 'd1' list."""
    box = rescale_frames(coords, dims)
    n, dr = grid_shape(box, sres, better_resolution)
    wrap_frames(coords, box)
    widths = half_widths(rad, dr, set(labels))
    nb = border_cells(widths)
    sf = np.zeros((int(n[0]), int(n[1]), int(n[2] / 2) + 1))
    if taps is not None:
        taps.update(N=n.copy(), dr=dr.copy(), B=nb, ir=[], d1=[])
    for t in range(coords.shape[0]):
        d1 = density_frame(coords[t], labels, rad, widths, n, dr, nb, ucell, fold_mode)
        if taps is not None:
            taps["ir"].append(cell_indices(coords[t], dr))
            taps["d1"].append(d1)
        sf += power_spectrum(d1)
        if frame_callback is not None:
            frame_callback(t)
    sfplt = centred_view(sf)
    kgrid, kplt = k_lattices(sf.shape, box)
    kplt[:, :, :, 3] = sfplt
    return dict(sf=sf, sfplt=sfplt, L=box, N=n, kgrid=kgrid, kgridplt=kplt)
