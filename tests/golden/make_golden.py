#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run once here (``python tests/golden/make_golden.py``); it needs ``/root/reference`` and is
never executed on the GPU box.  ``dens.py`` of the reference imports ``past.utils.old_div``
(python-future, not installed): a two-line shim with python-future's semantics is put on
sys.path for this process only.  The per-atom tqdm bars are silenced (no numeric effect).
Each fixture stores the inputs, the six arrays of the reference's sf npz, the in-place
mutated coordinates, and the periodic density d1 of every frame captured by wrapping
``np.fft.rfftn``.
"""
import math
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference_dens():
    shim = tempfile.mkdtemp(prefix="past_shim_")
    os.makedirs(os.path.join(shim, "past"))
    open(os.path.join(shim, "past", "__init__.py"), "w").close()
    with open(os.path.join(shim, "past", "utils.py"), "w") as fh:
        fh.write("import numbers\n"
                 "def old_div(a, b):\n"
                 "    if isinstance(a, numbers.Integral) and isinstance(b, numbers.Integral):\n"
                 "        return a // b\n"
                 "    return a / b\n")
    sys.path.insert(0, shim)
    sys.path.insert(0, REF)
    import tqdm
    passthrough = lambda it=None, *a, **k: it
    tqdm.tqdm = passthrough
    import dens
    dens.tqdm.tqdm = passthrough
    dens.trange = lambda *a, **k: range(*a)
    return dens


def run_reference(dens, coords, dims, typ, rad, ucell, sres):
    captured = []
    real_rfftn = np.fft.rfftn

    def spy(a, *args, **kw):
        captured.append(np.array(a, copy=True))
        return real_rfftn(a, *args, **kw)

    out = os.path.join(tempfile.mkdtemp(prefix="gold_"), "sf")
    np.fft.rfftn = spy
    try:
        dens.compute_sf(coords, dims, typ, out, rad, ucell, sres)
    finally:
        np.fft.rfftn = real_rfftn
    z = np.load(out + ".npz")
    return {k: z[k] for k in z.files}, np.stack(captured)


def ucell_for(theta_deg):
    th = theta_deg * math.pi / 180.0    # main_gromacs.py:79-80
    return np.array([[1, 0, 0], [np.cos(th), np.sin(th), 0], [0, 0, 1]])


def make_case(name, seed, nframes, natoms, box, jitter, labels, theta, sres, cdt, ddt, spread=(-0.3, 1.3),
              corner_bias=False):
    rng = np.random.default_rng(seed)
    dims = (np.asarray(box)[None, :] * (1.0 + jitter * rng.standard_normal((nframes, 3)))).astype(ddt)
    u = rng.uniform(spread[0], spread[1], size=(nframes, natoms, 3))
    if corner_bias:   # push half of the atoms next to the box corners/edges so every fold region is hit
        k = natoms // 2
        u[:, :k, :] = np.where(rng.random((nframes, k, 3)) < 0.5, 0.02, 0.97) + 0.02 * rng.random((nframes, k, 3))
    coords = (u * dims[:, None, :].astype(np.float64)).astype(cdt)
    typ = np.array([labels[i % len(labels)] for i in rng.permutation(natoms)])
    return dict(name=name, coords=coords, dims=dims, typ=typ, ucell=ucell_for(theta) if theta else np.eye(3),
                sres=sres)


def main():
    dens = import_reference_dens()
    rad = dens.load_radii(os.path.join(REF, "radii.txt"))
    cases = [
        make_case("gas_f64_ortho", 11, 2, 40, (12.0, 11.0, 10.0), 0.01, ["H", "C", "O"], None, 1.0,
                  np.float64, np.float64),
        make_case("mono_f32", 12, 3, 60, (14.0, 14.0, 12.5), 0.01, ["C", "H", "O", "N", "H30", "C14"], 120.0, 0.9,
                  np.float32, np.float32),
        make_case("corner_na_f64", 13, 1, 24, (20.0, 20.0, 19.0), 0.0, ["NA", "C", "H", "O"], 120.0, 1.0,
                  np.float64, np.float64, spread=(0.0, 1.0), corner_bias=True),
        make_case("mixed_f32c_f64d", 14, 2, 30, (10.0, 12.0, 11.0), 0.02, ["C", "H"], 120.0, 1.0,
                  np.float32, np.float64),
        make_case("mixed_f64c_f32d", 15, 2, 30, (10.0, 12.0, 11.0), 0.02, ["O", "H"], 60.0, 1.0,
                  np.float64, np.float32),
        make_case("fine_f32", 16, 1, 12, (8.0, 7.0, 9.0), 0.0, ["H", "O"], 120.0, 0.5,
                  np.float32, np.float32, spread=(0.0, 1.0), corner_bias=True),
        make_case("odd_sizes_f32", 17, 2, 50, (13.3, 10.7, 9.1), 0.005, ["C", "H", "O", "N"], 120.0, 1.0,
                  np.float32, np.float32),
    ]
    for c in cases:
        r_in = c["coords"].copy()
        r = c["coords"].copy()
        labels = sorted(set(c["typ"].tolist()))
        out, d1 = run_reference(dens, r, c["dims"].copy(), c["typ"], rad, c["ucell"], c["sres"])
        path = os.path.join(HERE, c["name"] + ".npz")
        np.savez_compressed(
            path, coords=r_in, dims=c["dims"], typ=c["typ"], ucell=c["ucell"], sres=np.float64(c["sres"]),
            rad_labels=np.array(labels), rad_nel=np.array([rad[l][0] for l in labels]),
            rad_sigma=np.array([rad[l][1] for l in labels]),
            coords_after=r, d1=d1, **{"ref_" + k: v for k, v in out.items()})
        print(c["name"], "N=", out["N"], "d1", d1.shape, "%.1f kB" % (os.path.getsize(path) / 1e3))

    # stand-alone helpers of the reference, on random inputs
    rng = np.random.default_rng(99)
    fold_cases = {}
    for i, (n, b) in enumerate([((12, 10, 14), 3), ((8, 8, 8), 4), ((20, 16, 12), 5), ((6, 10, 8), 6)]):
        d0 = rng.random((n[0] + 2 * b, n[1] + 2 * b, n[2] + 2 * b))
        des = [[0, b, n[d] - b, n[d]] for d in range(3)]
        ori = [[0, b, b + n[d], 2 * b + n[d]] for d in range(3)]
        fold_cases["fold%d_d0" % i] = d0
        fold_cases["fold%d_nb" % i] = np.array(list(n) + [b])
        fold_cases["fold%d_d1" % i] = dens.remap_grid_tcl(d0, des, ori)
    for i, shp in enumerate([(8, 6, 5), (12, 12, 7), (10, 14, 9)]):
        sf = rng.random(shp)
        fold_cases["dplot%d_in" % i] = sf
        fold_cases["dplot%d_out" % i] = dens.get_dplot(sf)
    np.savez_compressed(os.path.join(HERE, "helpers.npz"), **fold_cases)
    print("helpers ok")


if __name__ == "__main__":
    main()
