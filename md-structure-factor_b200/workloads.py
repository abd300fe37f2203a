"""Synthetic trajectories for the benchmark configurations of BASELINE.json.

All coordinates are float32 Angstrom and lie strictly inside (0, L), so the reference's PBC pass
(and its >= 1e6-atom quirk, dens.py:213-221) is a no-op on them.  Generators are deterministic
(numpy default_rng with the stated seeds).  Pure numpy; used by bench.py and the tests.
"""
import math

import numpy as np

# (electrons, sigma in Angstrom) of the labels the workloads use -- values as in radii.txt
RAD = {"H": (1.0, 0.53), "C": (6.0, 0.70), "N": (7.0, 0.56), "O": (8.0, 0.60), "NA": (11.0, 2.27)}


def ucell_for(theta_deg):
    """ucell rows for cell angle theta (reference main_gromacs.py:79-80)."""
    th = theta_deg * math.pi / 180.0
    return np.array([[1, 0, 0], [np.cos(th), np.sin(th), 0], [0, 0, 1]])


def _inside(x, box):
    eps = np.float32(1e-3)
    x = np.mod(x, box).astype(np.float32)
    return np.clip(x, eps, (box - eps).astype(np.float32))


def jitter_frames(base, box, nframes, sigma, seed0, out=None):
    """frame t = base + N(0, sigma), seed seed0+t, wrapped strictly inside the box."""
    out = np.empty((nframes,) + base.shape, dtype=np.float32) if out is None else out
    for t in range(nframes):
        rng = np.random.default_rng(seed0 + t)
        out[t] = _inside(base + rng.normal(0.0, sigma, size=base.shape).astype(np.float32), box)
    return out


def water_box(reps=26):
    """c2: a 2-molecule SPC water cell (18.206 A cubic, the size of the reference's test/water.gro)
    replicated reps^3 times: 6*reps^3 atoms labelled O,H,H.  reps=26 -> 105 456 atoms, L=473.356 A."""
    cell = 18.206
    oh, ang = 1.0, math.radians(109.47)
    mol = np.array([[0.0, 0.0, 0.0], [oh, 0.0, 0.0], [oh * math.cos(ang), oh * math.sin(ang), 0.0]])
    rot = np.array([[0.36, -0.48, 0.8], [0.8, 0.6, 0.0], [-0.48, 0.64, 0.6]])
    unit = np.concatenate([mol + np.array([1.26, 16.24, 16.79]), mol @ rot.T + np.array([12.75, 0.53, 6.22])])
    shifts = np.stack(np.meshgrid(*[np.arange(reps)] * 3, indexing="ij"), -1).reshape(-1, 1, 3) * cell
    base = (unit[None] + shifts).reshape(-1, 3).astype(np.float32)
    box = np.full(3, cell * reps, dtype=np.float32)
    typ = np.array(["O", "H", "H"] * (2 * reps ** 3))
    return typ, _inside(base, box), box


def uniform_box(natoms, L, composition, seed):
    """Uniform random atoms in a cubic box with the given label fractions (c4/c5 recipe)."""
    rng = np.random.default_rng(seed)
    box = np.full(3, L, dtype=np.float32)
    base = _inside(rng.uniform(0.0, L, size=(natoms, 3)).astype(np.float32), box)
    labels, frac = zip(*composition)
    typ = rng.choice(np.array(labels), size=natoms, p=np.array(frac) / sum(frac))
    return typ, base, box


def _test_system():
    """c1 as specified: frame 0 = the reference's test/test_system.gro (a gzip copy travels as
    tests/golden/test_system.gro.gz: 55 680 atoms, hexagonal box, atom NAMES as labels, sodium stamps of 34^3 cells), transformed
    like main_gromacs.py:204-207 (theta = 120), Sres = 1 -> 88 x 88 x 84; frames = frame 0 + N(0, 0.3 A).  None when the fixture
    is not there (the xtc of the reference is a missing blob anyway)."""
    import gzip
    import os
    import shutil
    import tempfile
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, os.pardir, "tests", "golden", "test_system.gro.gz")
    if not os.path.exists(path):
        return None
    import dens
    import load_traj
    with tempfile.TemporaryDirectory() as tmp:
        gro = os.path.join(tmp, "test_system.gro")
        with gzip.open(path, "rb") as src, open(gro, "wb") as dst:
            shutil.copyfileobj(src, dst)
        names, coords, dims = load_traj.read_gro_frames(gro)
    theta = 120.0 * math.pi / 180.0
    base = coords[0].copy()
    base[:, 1] = base[:, 1] / np.sin(theta)
    base[:, 0] = base[:, 0] - base[:, 1] * np.cos(theta)
    box = dims[0].astype(np.float32)
    rad = dens.load_radii(os.path.join(here, "radii.txt"))
    return dict(typ=np.array(names), base=_inside(base, box), box=box, ucell=ucell_for(120.0), sres=1.0, grid=(88, 88, 84), jitter=0.3,
                seed0=1000, rad=rad, desc="test/test_system.gro of the reference (55680 atoms, hexagonal box), monoclinic transform, 88x88x84, theta=120")


LLC_COMPOSITION = [("C", 0.35), ("H", 0.58), ("O", 0.06), ("N", 0.003), ("NA", 0.007)]
BOX_COMPOSITION = [("C", 0.40), ("H", 0.50), ("O", 0.08), ("N", 0.01), ("NA", 0.01)]


def get(name):
    """Workload by config id -> dict(typ, base, box, ucell, sres, grid, jitter, seed0, rad, desc)."""
    if name == "c2":
        typ, base, box = water_box(26)
        n = 256
        return dict(typ=typ, base=base, box=box, ucell=np.eye(3), sres=float(box[0]) / n * (1 + 1e-6), grid=(n, n, n),
                    jitter=0.5, seed0=2000, rad=RAD, desc="SPC water 26^3 replicas, 105456 atoms, 256^3 grid, ucell=I")
    if name == "c1":   # BASELINE configs[0]: the reference's own fixture test/test_system.gro through the CLI's monoclinic transform
        real = _test_system()
        if real is not None:
            return real
        name = "c1u"
    if name == "c1u":  # size/composition of test_system.gro (55 680 atoms, hexagonal 86.59 x 86.59 x 83.43 A), theta=120, uniform random
        rng = np.random.default_rng(1000)
        box = np.array([86.5915, 86.591576, 83.431305], dtype=np.float32)
        natoms = 55680
        base = _inside((rng.uniform(0, 1, size=(natoms, 3)) * box).astype(np.float32), box)
        labels, frac = zip(*LLC_COMPOSITION)
        typ = rng.choice(np.array(labels), size=natoms, p=np.array(frac) / sum(frac))
        return dict(typ=typ, base=base, box=box, ucell=ucell_for(120.0), sres=1.0, grid=(88, 88, 84), jitter=0.3,
                    seed0=1000, rad=RAD, desc="LLC-composition box of test_system.gro's size, 55680 atoms, 88x88x84, theta=120")
    if name == "c3":   # LLC composition, 512^3 grid at the reference's default 1 A resolution... see DESIGN.md
        typ, base, box = uniform_box(334080, 511.9, LLC_COMPOSITION, 3000)
        return dict(typ=typ, base=base, box=box, ucell=ucell_for(120.0), sres=1.0, grid=(512, 512, 512), jitter=0.3,
                    seed0=3000, rad=RAD, desc="LLC-composition box, 334080 atoms, 512^3 grid, theta=120")
    if name == "c3s":  # development aid: c3's grid with a tenth of its atoms (isolates the grid passes from the atom work)
        typ, base, box = uniform_box(33408, 511.9, LLC_COMPOSITION, 3000)
        return dict(typ=typ, base=base, box=box, ucell=ucell_for(120.0), sres=1.0, grid=(512, 512, 512), jitter=0.3,
                    seed0=3000, rad=RAD, desc="c3 grid, 33408 atoms")
    if name == "c3d":  # c3's atoms at the number density of a condensed phase (0.1 atoms / A^3, what an LLC membrane or water has):
        # the 512^3 grid then resolves 0.29 A, stamps are 26^3 (H) ... 36^3 (C) ... 118^3 (NA) cells and the splat dominates
        L = (334080 / 0.1) ** (1.0 / 3.0)
        typ, base, box = uniform_box(334080, L, LLC_COMPOSITION, 3100)
        return dict(typ=typ, base=base, box=box, ucell=ucell_for(120.0), sres=float(box[0]) / 512 * (1 + 1e-6), grid=(512, 512, 512),
                    jitter=0.3, seed0=3100, rad=RAD, desc="LLC-composition box at 0.1 atoms/A^3, 334080 atoms, 512^3 grid (dr = 0.29 A), theta=120")
    if name == "c4":
        typ, base, box = uniform_box(1000000, 767.9, BOX_COMPOSITION, 4000)
        return dict(typ=typ, base=base, box=box, ucell=np.eye(3), sres=1.0, grid=(768, 768, 768), jitter=0.3,
                    seed0=4000, rad=RAD, desc="1M-atom uniform box, 768^3 grid, per-element form factors")
    if name == "c5":
        typ, base, box = uniform_box(4000000, 1023.9, BOX_COMPOSITION, 5000)
        return dict(typ=typ, base=base, box=box, ucell=np.eye(3), sres=1.0, grid=(1024, 1024, 1024), jitter=0.3,
                    seed0=5000, rad=RAD, desc="4M-atom uniform box, 1024^3 grid")
    if name == "tiny":
        typ, base, box = water_box(3)
        return dict(typ=typ, base=base, box=box, ucell=ucell_for(120.0), sres=1.0, grid=(56, 56, 56), jitter=0.5,
                    seed0=10, rad=RAD, desc="162-atom water box, smoke test")
    raise KeyError(name)


def algorithmic_bytes_per_frame(grid, natoms, coord_bytes=4):
    """Na*(3*sizeof(coord)+4) + 8 N^3 (density write) + 8 N^3 (density read by the FFT)
    + 2*8*Nx*Ny*(Nz/2+1) (S(q) read-modify-write)."""
    nx, ny, nz = grid
    return natoms * (3 * coord_bytes + 4) + 16 * nx * ny * nz + 16 * nx * ny * (nz // 2 + 1)
