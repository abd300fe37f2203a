"""CPU tests of the host side of the drop-in boundary (no GPU, no compute calls): the numpy parts
of dens.py against the reference's golden outputs, the C ABI surface, the loaders, workloads."""
import ctypes
import os
import re

import numpy as np
import pytest

import mdsf_b200
from tests.helpers import CASES, GOLDEN, ROOT, load_case

dens = mdsf_b200.dens


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mdsf.h")).read()
    declared = sorted(set(re.findall(r"\b(mdsf_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 20
    lib = mdsf_b200.native.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(mdsf_b200.native.EXPORTS) == declared
    assert lib.mdsf_abi_version() == mdsf_b200.native.ABI_VERSION


def test_config_struct_matches_header_layout():
    # 2 + 3 + 1 int32 (24 B) | 3+3+9 doubles (120 B) | int32 + pad | 3 pointers | 8 int32 | 7 reserved
    cfg = mdsf_b200.native.Config
    assert cfg.dr.offset == 24 and cfg.box.offset == 48 and cfg.ucell.offset == 72
    assert cfg.ntypes.offset == 144 and cfg.amp.offset == 152 and cfg.halfw.offset == 168
    assert cfg.coord_dtype.offset == 176 and ctypes.sizeof(cfg) == 240


def test_no_cuda_device_fails_loudly_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mdsf_b200.native.MdsfError) as ei:
        mdsf_b200.native.Engine((8, 8, 8), 2, (1, 1, 1), (8, 8, 8), np.eye(3), [1.0], [1.0], [[2, 2, 2]], np.float32, np.float32)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


@pytest.mark.parametrize("name", CASES)
def test_grid_borders_and_k_lattices_match_reference(name):
    c = load_case(name)
    L = np.average(c["dims"], axis=0)
    assert L.dtype == c["ref_L"].dtype and np.array_equal(L, c["ref_L"])
    n, dr = dens._grid(L, c["sres"])
    assert np.array_equal(n, c["ref_N"]) and dr.dtype == np.float64
    kgrid, kplt = dens._k_lattices(c["ref_sf"].shape, L)
    assert np.array_equal(kgrid, c["ref_kgrid"])
    assert np.array_equal(kplt[..., :3], c["ref_kgridplt"][..., :3])
    assert np.array_equal(dens.get_dplot(c["ref_sf"]), c["ref_sfplt"])


def test_remap_grid_tcl_and_get_dplot_match_reference_helpers():
    z = np.load(os.path.join(GOLDEN, "helpers.npz"))
    i = 0
    while "fold%d_d0" % i in z.files:
        n = z["fold%d_nb" % i]
        b = int(n[3])
        des = [[0, b, int(n[d]) - b, int(n[d])] for d in range(3)]
        ori = [[0, b, b + int(n[d]), 2 * b + int(n[d])] for d in range(3)]
        assert np.abs(dens.remap_grid_tcl(z["fold%d_d0" % i], des, ori) - z["fold%d_d1" % i]).max() < 1e-13
        i += 1
    i = 0
    while "dplot%d_in" % i in z.files:
        assert np.array_equal(dens.get_dplot(z["dplot%d_in" % i]), z["dplot%d_out" % i])
        i += 1


def test_rescale_helper_is_in_place_and_matches_golden():
    c = load_case("mono_f32")
    r = c["coords"].copy()
    out, L = dens.rescale(r, c["dims"])
    assert out is r and np.array_equal(L, c["ref_L"])
    from oracle import dens_oracle as orc
    r2 = c["coords"].copy()
    orc.rescale_frames(r2, c["dims"])
    assert np.array_equal(r, r2)


def test_radii_table_is_a_superset_with_reference_values(tmp_path):
    rad = dens.load_radii(os.path.join(ROOT, "md-structure-factor_b200", "radii.txt"))
    for lab, val in {"H": (1.0, 0.53), "C": (6.0, 0.70), "O": (8.0, 0.60), "N": (7.0, 0.56), "NA": (11.0, 2.27),
                     "R3": (1.0, 0.90), "C145": (6.0, 0.67), "OES": (8.0, 0.48), "H146": (1.0, 0.53)}.items():
        assert rad[lab] == val
    for c in CASES:   # every label the golden cases use has the value the reference's table gave
        for lab, val in load_case(c)["rad"].items():
            assert rad[lab] == val
    p = tmp_path / "r.txt"
    p.write_text("1 H 53\n8\tO\t60\n1 H 99\n")
    assert dens.load_radii(str(p)) == {"H": (1.0, 0.99), "O": (8.0, 0.60)}     # later rows win
    assert set(dens.get_borders({"H": (1.0, 0.53)}, np.ones(3), ["H"]).keys()) == {"H"}


def test_large_system_wrap_quirk_range():
    assert dens._wrapped_atoms(5, 999999) == (0, 999999)
    assert dens._wrapped_atoms(5, 1000000) == (0, 5)
    assert dens._wrapped_atoms(3000000, 4000000) == (0, 0)          # imax == nframes: the branch never fires
    assert dens._wrapped_atoms(2500000, 4000000) == (2000000, 2500000)


def test_module_knobs_exist_like_the_reference():
    for attr in ("USE_BETTER_RESOLUTION", "PRINT_DETAILS", "RANDOM_NOISE", "Nspatialgrid", "theta", "PRECISION", "dtyp",
                 "load_radii", "get_borders", "rescale", "remap_grid_tcl", "get_dplot", "compute_sf"):
        assert hasattr(dens, attr)
    assert dens.PRECISION == 1.0e-24 and dens.dtyp is np.float64


def test_load_gro_and_traj_npz_layout(tmp_path):
    lt = mdsf_b200.load_traj
    gro = tmp_path / "w.gro"
    gro.write_text("two waters\n    6\n"
                   "    1WATER  OW1    1   0.126   1.624   1.679\n    1WATER  HW2    2   0.190   1.661   1.747\n"
                   "    1WATER  HW3    3   0.177   1.568   1.613\n    2WATER  OW1    4   1.275   0.053   0.622\n"
                   "    2WATER  HW2    5   1.337   0.002   0.680\n    2WATER  HW3    6   1.326   0.120   0.568\n"
                   "   1.82060   1.82060   1.82060\n")
    assert lt.load_gro(str(gro)) == ["OW1", "HW2", "HW3", "OW1", "HW2", "HW3"]     # the reference's own unit test
    names, xyz, box = lt.read_gro(str(gro))
    assert xyz.dtype == np.float32 and xyz.shape == (6, 3) and np.allclose(box, 18.206)
    assert np.allclose(xyz[0], [1.26, 16.24, 16.79])
    lt.process_gro_mdtraj(str(gro), str(gro), str(tmp_path / "out_w_traj"))
    z = np.load(str(tmp_path / "out_w_traj.npz"))
    assert sorted(z.files) == ["coords", "dims", "mass", "name", "typ"]
    assert z["coords"].shape == (1, 6, 3) and z["coords"].dtype == np.float32 and z["dims"].shape == (1, 3)
    assert list(z["typ"]) == names                                                  # typ = atom NAMES (load_traj.py:110)
    # higher-precision files (gmx ... -ndec 5 writes %10.5f fields): the width follows the decimal points, as in GROMACS / mdtraj
    hp = tmp_path / "hp.gro"
    hp.write_text("hp\n    2\n"
                  "    1WATER  OW1    1   0.12600   1.62400   1.67900\n    1WATER  HW2    2  -0.19012  11.66123   1.74700\n"
                  "   1.82060   1.82060   1.82060\n")
    _, xyz5, _ = lt.read_gro(str(hp))
    assert np.allclose(xyz5, [[1.26, 16.24, 16.79], [-1.9012, 116.6123, 17.47]])
    _, xyz5f, _ = lt.read_gro_frames(str(hp))
    assert np.array_equal(xyz5f[0], xyz5)


def test_c1_workload_is_the_reference_fixture():
    """BASELINE configs[0]: workloads.get('c1') is the reference's test/test_system.gro frame (gzip copy under tests/golden),
    monoclinic-transformed like main_gromacs.py:204-207, labelled with the atom NAMES (load_traj.py:110), 88 x 88 x 84 at Sres = 1."""
    w = __import__("workloads")
    c = w.get("c1")
    assert c["base"].shape == (55680, 3) and c["base"].dtype == np.float32 and c["grid"] == (88, 88, 84)
    assert "test_system.gro" in c["desc"] and len(set(c["typ"])) > 100 and "NA" in set(c["typ"])
    assert all(label in c["rad"] for label in set(c["typ"]))            # every name has a form factor (radii.txt superset)
    n, dr = mdsf_b200.dens._grid(np.asarray(c["box"]), c["sres"])
    assert tuple(int(v) for v in n) == c["grid"]
    assert np.all(c["base"] > 0) and np.all(c["base"] < c["box"])
    u = w.get("c1u")
    assert u["base"].shape == (55680, 3) and "LLC-composition" in u["desc"]


def test_workloads_are_deterministic_and_inside_the_box():
    w = __import__("workloads")
    c = w.get("tiny")
    a = w.jitter_frames(c["base"], c["box"], 2, c["jitter"], c["seed0"])
    b = w.jitter_frames(c["base"], c["box"], 2, c["jitter"], c["seed0"])
    assert np.array_equal(a, b) and a.dtype == np.float32
    assert (a > 0).all() and (a < c["box"]).all()
    c2 = w.get("c2")
    assert c2["base"].shape == (105456, 3) and tuple(dens._grid(c2["box"], c2["sres"])[0]) == (256, 256, 256)
    assert w.algorithmic_bytes_per_frame((256, 256, 256), 105456) == 105456 * 16 + 16 * 256 ** 3 + 16 * 256 * 256 * 129


def test_parallel_npz_writer_matches_numpy_container(tmp_path):
    """npz_writer.savez_parallel writes what np.savez_compressed writes (reference dens.py:346): same keys, dtypes,
    shapes and values through np.load, valid CRCs, also with many small chunks, zip64 records, Fortran order, 0-d,
    empty and stored members."""
    import zipfile
    import npz_writer
    rng = np.random.default_rng(5)
    arrays = dict(sf=rng.random((12, 10, 7)), sfplt=rng.random((10, 8, 13)), L=np.array([1.5, 2.5, 3.5], dtype=np.float32),
                  N=np.array([12, 10, 12]), kgrid=np.zeros((12, 10, 7, 4)), kgridplt=np.asfortranarray(rng.random((10, 8, 13, 4))),
                  scalar=np.float64(3.25), empty=np.zeros((0, 3)), names=np.array(["OW", "HW1", "HW2"]))
    ref = str(tmp_path / "ref.npz")
    np.savez_compressed(ref, **arrays)
    want = np.load(ref)
    for kw in (dict(), dict(chunk=1000, threads=3), dict(chunk=777, force_zip64=True), dict(compressed=False), dict(compressed=False, force_zip64=True)):
        out = npz_writer.savez_parallel(str(tmp_path / "par"), **kw, **arrays)
        assert out.endswith("par.npz")
        with zipfile.ZipFile(out) as zf:
            assert zf.testzip() is None
            assert sorted(zf.namelist()) == sorted(k + ".npy" for k in arrays)
            assert all(i.compress_type == (zipfile.ZIP_STORED if kw.get("compressed") is False else zipfile.ZIP_DEFLATED) for i in zf.infolist())
        got = np.load(out)
        assert sorted(got.files) == sorted(want.files)
        for k in want.files:
            assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), k
            assert got[k].flags.f_contiguous == want[k].flags.f_contiguous


def test_npz_frame_stream_reads_chunks_like_np_load(tmp_path):
    """load_traj.NpzFrameStream inflates the coords member of a traj npz (reference load_traj.py:110) incrementally."""
    import load_traj
    import npz_writer
    rng = np.random.default_rng(9)
    coords = rng.normal(size=(7, 11, 3)).astype(np.float32)
    dims = np.full((7, 3), 20.0, dtype=np.float32)
    names = np.array(["C%d" % i for i in range(11)])
    a = str(tmp_path / "a_traj")
    load_traj.save_traj_npz(a, dims, coords, names)
    b = npz_writer.savez_parallel(str(tmp_path / "b_traj"), chunk=100, dims=dims, coords=coords, name=names, mass=np.zeros(11), typ=names)
    for path in (a + ".npz", b):
        with load_traj.NpzFrameStream(path) as fs:
            assert fs.shape == coords.shape and fs.dtype == coords.dtype
            fs.skip(1)
            buf = np.zeros((4, 11, 3), dtype=np.float32)
            assert fs.read_into(buf) == 4 and np.array_equal(buf, coords[1:5])
            assert fs.read_into(buf) == 2 and np.array_equal(buf[:2], coords[5:7])
            assert fs.read_into(buf) == 0
            with pytest.raises(ValueError):
                fs.read_into(np.zeros((4, 11, 3), dtype=np.float64))


def test_io_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mdsf_io.h")).read()
    declared = sorted(set(re.findall(r"\b(mdsf_io_[a-z_0-9]+)\s*\(", header)))
    import load_traj
    lib = load_traj._io_lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(load_traj.IO_EXPORTS) == declared
    assert lib.mdsf_io_abi_version() == 2


def test_indexed_traj_npz_is_inflated_in_parallel_and_stays_a_numpy_container(tmp_path):
    """save_traj_npz writes the reference's container (load_traj.py:110) as a chain of independently deflated pieces
    plus a piece index in a private zip extra field; NpzFrameStream inflates those pieces on all cores through
    libmdsf_io (straight into the chunk buffer) and falls back to sequential inflate for plain np.savez_compressed files."""
    import zipfile
    import load_traj
    import npz_writer
    rng = np.random.default_rng(5)
    T, na = 29, 4001
    coords = rng.uniform(0, 60, size=(T, na, 3)).astype(np.float32)
    dims = np.full((T, 3), 60.0, dtype=np.float32)
    names = np.array(["O", "H", "H", "C", "NA", "N", "C"] * (na // 7 + 1))[:na]
    path = npz_writer.savez_parallel(str(tmp_path / "p_traj"), chunk=70001, dims=dims, coords=coords, name=names, mass=np.zeros(na), typ=names)
    z = np.load(path)                                         # numpy and the reference read it as before
    assert z.files == ["dims", "coords", "name", "mass", "typ"] and np.array_equal(z["coords"], coords) and list(z["typ"]) == list(names)
    assert zipfile.ZipFile(path).testzip() is None            # CRCs and sizes of every member are right
    ix = npz_writer.read_piece_index(path, "coords.npy")
    assert ix is not None and len(ix["raw_len"]) == -(-(coords.nbytes + 128) // 70001) and sum(ix["raw_len"]) == ix["raw_size"]
    assert npz_writer.read_piece_index(path, "dims.npy") is None          # single-piece members carry no index
    for chunk_frames, threads in ((1, 1), (4, 3), (9, 0), (64, 0)):
        with load_traj.NpzFrameStream(path, threads=threads) as fs:
            assert fs.parallel and fs.shape == coords.shape
            fs.skip(2)
            got, buf = [], np.zeros((chunk_frames, na, 3), dtype=np.float32)
            while True:
                k = fs.read_into(buf)
                if k == 0:
                    break
                got.append(buf[:k].copy())
            assert np.array_equal(np.concatenate(got), coords[2:]), (chunk_frames, threads)
    ref = str(tmp_path / "r_traj.npz")
    np.savez_compressed(ref, dims=dims, coords=coords, name=names, mass=np.zeros(na), typ=names)
    with load_traj.NpzFrameStream(ref) as fs:
        assert not fs.parallel
        buf = np.zeros((T, na, 3), dtype=np.float32)
        assert fs.read_into(buf) == T and np.array_equal(buf, coords)
    # a piece whose bytes disagree with the index is reported, not silently mis-decoded
    blob = bytearray(open(path, "rb").read())
    mid = ix["data_offset"] + ix["comp_start"][3] + 4
    blob[mid:mid + 64] = b"\x00" * 64
    bad = str(tmp_path / "bad_traj.npz")
    open(bad, "wb").write(bytes(blob))
    with load_traj.NpzFrameStream(bad) as fs, pytest.raises(EOFError):
        fs.read_into(np.zeros((T, na, 3), dtype=np.float32))


def test_native_trr_and_multiframe_gro_readers(tmp_path):
    """load_traj.read_trr follows the published .trr layout (magic 1993, "GMX_trn_file", block sizes, reals in nm) for float
    and double files, skips coordinate-free frames and feeds process_gro_mdtraj when mdtraj is absent; read_gro_frames
    splits concatenated .gro frames.  (Parity unpinned: no mdtraj / fixture here, see the reader's docstring.)"""
    import load_traj as lt
    rng = np.random.default_rng(3)
    T, na = 5, 6
    x = rng.uniform(0, 1.8, size=(T, na, 3))
    box = np.array([[1.8206, 1.8206 * (1 + 0.01 * t), 1.75] for t in range(T)])
    for dbl in (False, True):
        p = str(tmp_path / ("d.trr" if dbl else "f.trr"))
        lt.write_trr(p, x, box, times=np.arange(T) * 2.0, double=dbl, velocities=x * 0.1)
        c, b, t = lt.read_trr(p)
        assert c.dtype == np.float32 and c.shape == (T, na, 3) and b.dtype == np.float32
        ref = x.astype(np.float64 if dbl else np.float32).astype(np.float32) * np.float32(10)
        assert np.array_equal(c, ref) and np.allclose(b, box * 10, rtol=1e-6) and np.array_equal(t, np.arange(T) * 2.0)
    raw = open(str(tmp_path / "f.trr"), "rb").read()
    assert raw[:4] == b"\x00\x00\x07\xc9" and raw[4:12] == b"\x00\x00\x00\x0d\x00\x00\x00\x0c" and raw[12:24] == b"GMX_trn_file"
    tri = np.array([[[2.0, 0, 0], [-1.0, 1.7320508, 0], [0, 0, 3.0]]] * T)          # hexagonal box vectors -> lengths 2, 2, 3
    lt.write_trr(str(tmp_path / "h.trr"), x, tri)
    assert np.allclose(lt.read_trr(str(tmp_path / "h.trr"))[1], [20.0, 20.0, 30.0], rtol=1e-6)
    with pytest.raises(ValueError):
        open(str(tmp_path / "bad.trr"), "wb").write(raw[:100])
        lt.read_trr(str(tmp_path / "bad.trr"))
    # topology .gro + .trr -> traj npz without mdtraj
    gro = tmp_path / "w.gro"
    rows = ["    1WATER%5s%5d%8.3f%8.3f%8.3f\n" % (nm, i + 1, *x[0, i]) for i, nm in enumerate(["OW1", "HW2", "HW3"] * 2)]
    gro.write_text("two waters\n    6\n" + "".join(rows) + "   1.82060   1.82060   1.75000\n")
    try:
        import mdtraj  # noqa: F401
    except ImportError:
        lt.process_gro_mdtraj(str(gro), str(tmp_path / "f.trr"), str(tmp_path / "out_f_traj"))
        z = np.load(str(tmp_path / "out_f_traj.npz"))
        assert z["coords"].shape == (T, na, 3) and z["dims"].shape == (T, 3) and list(z["typ"]) == ["OW1", "HW2", "HW3"] * 2
        assert np.array_equal(z["coords"], x.astype(np.float32) * np.float32(10))
    # concatenated .gro frames
    multi = tmp_path / "m.gro"
    multi.write_text(gro.read_text() * 3)
    names, c, b = lt.read_gro_frames(str(multi))
    assert c.shape == (3, na, 3) and b.shape == (3, 3) and names == ["OW1", "HW2", "HW3"] * 2
    assert np.array_equal(c[0], c[2]) and np.allclose(b[1], [18.206, 18.206, 17.5])


def _xtc_expected(x_nm, precision):
    """what an xtc3 round trip does to float32 nm coordinates: round(x * precision) / precision in float32, then Angstrom"""
    v = x_nm.astype(np.float64) * np.float32(precision)
    q = np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int32)
    return q.astype(np.float32) * (np.float32(1) / np.float32(precision)) * np.float32(10)


def test_native_xtc_reader_round_trips_every_coding_mode(tmp_path):
    """load_traj.read_xtc (headers in Python, xtc3 blocks in libmdsf_io on host threads) against load_traj.write_xtc, a
    separate restatement of the encoder: water (runs of small offsets with the leader swap), a random walk (long runs,
    small-range adaptation both ways), a gas (no runs), a box wider than 24 bits of precision units (plain bit fields
    instead of the mixed-radix triple), negative coordinates, <= 9 atoms (plain floats).  Bit-exact on the integers,
    i.e. on the float32 output.  (Parity unpinned: no mdtraj / .xtc fixture here, see the reader's docstring.)"""
    import load_traj as lt
    rng = np.random.default_rng(11)
    no = 300
    O = rng.uniform(0, 3, size=(3, no, 3))
    water = np.stack([O, O + rng.normal(0, 0.05, O.shape), O + rng.normal(0, 0.05, O.shape)], axis=2).reshape(3, no * 3, 3)
    cases = {
        "water": (water.astype(np.float32), 1000.0),
        "walk": ((np.cumsum(rng.normal(0, 0.03, size=(2, 1500, 3)), axis=1) + 5).astype(np.float32), 1000.0),
        "gas": (rng.uniform(-2, 12, size=(2, 700, 3)).astype(np.float32), 1000.0),
        "wide": (rng.uniform(0, 30, size=(2, 40, 3)).astype(np.float32), 1e6),
        "coarse": (rng.uniform(0, 4, size=(2, 64, 3)).astype(np.float32), 100.0),
        "ten": (rng.uniform(0, 2, size=(4, 10, 3)).astype(np.float32), 1000.0),
    }
    for tag, (x, prec) in cases.items():
        T = x.shape[0]
        p = str(tmp_path / (tag + ".xtc"))
        box = np.array([[3.0, 3.0 * (1 + 0.01 * t), 2.5] for t in range(T)])
        lt.write_xtc(p, x, box, times=np.arange(T) * 0.5, precision=prec)
        for threads in (1, 0):
            c, b, t = lt.read_xtc(p, threads=threads)
            assert c.dtype == np.float32 and c.shape == x.shape and b.dtype == np.float32, tag
            assert np.array_equal(c, _xtc_expected(x, prec)), tag
            assert np.allclose(b, box * 10, rtol=1e-6) and np.array_equal(t, np.arange(T) * 0.5), tag
    # the same files as a frame source for dens.compute_sf_stream: chunked reads, skip, dims up front, same bits
    full, dims, _ = lt.read_xtc(str(tmp_path / "water.xtc"))
    with lt.XtcFrameStream(str(tmp_path / "water.xtc")) as fs:
        assert fs.shape == full.shape and fs.dtype == np.float32 and np.array_equal(fs.dims, dims) and fs.dims.dtype == np.float32
        buf = np.zeros((2,) + full.shape[1:], dtype=np.float32)
        fs.skip(1)
        assert fs.read_into(buf) == 2 and np.array_equal(buf, full[1:3])
        assert fs.read_into(buf) == 0
        with pytest.raises(ValueError):
            fs.read_into(np.zeros((2,) + full.shape[1:], dtype=np.float64))
    # water at precision 1000 packs to 4-5 bytes per atom, as the scheme is known to (a plain-float frame takes 12)
    assert 3.5 < os.path.getsize(str(tmp_path / "water.xtc")) / (3 * no * 3) < 5.5
    # <= 9 atoms: plain floats, bit-exact
    tiny = rng.uniform(0, 3, size=(2, 5, 3)).astype(np.float32)
    lt.write_xtc(str(tmp_path / "tiny.xtc"), tiny, np.full((2, 3), 3.0))
    assert np.array_equal(lt.read_xtc(str(tmp_path / "tiny.xtc"))[0], tiny * np.float32(10))
    raw = open(str(tmp_path / "water.xtc"), "rb").read()
    assert raw[:4] == b"\x00\x00\x07\xcb" and raw[4:8] == (3 * no).to_bytes(4, "big")
    # truncated file, corrupt header fields and a payload cut short are errors, not garbage coordinates
    for name, blob in (("cut.xtc", raw[:len(raw) // 2]), ("magic.xtc", b"\x00\x00\x07\xc9" + raw[4:])):
        open(str(tmp_path / name), "wb").write(blob)
        with pytest.raises(ValueError):
            lt.read_xtc(str(tmp_path / name))
    bad = bytearray(raw)
    bad[52 + 32:52 + 36] = (200).to_bytes(4, "big")          # smallidx outside the table
    open(str(tmp_path / "idx.xtc"), "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        lt.read_xtc(str(tmp_path / "idx.xtc"))
    # random damage anywhere in the file: an error or some coordinates, never a crash or an out-of-bounds write
    frng = np.random.default_rng(5)
    for trial in range(150):
        blob = bytearray(raw)
        for _ in range(int(frng.integers(1, 4))):
            at = int(frng.integers(0, 200)) if trial % 3 == 0 else int(frng.integers(0, len(blob)))
            blob[at] = int(frng.integers(0, 256))
        open(str(tmp_path / "fuzz.xtc"), "wb").write(bytes(blob))
        try:
            got = lt.read_xtc(str(tmp_path / "fuzz.xtc"), threads=1)[0]
            assert got.shape[1:] == (3 * no, 3)
        except ValueError:
            pass
    # topology .gro + .xtc -> traj npz without mdtraj (reference load_traj.py:90-111)
    x = cases["ten"][0]
    gro = tmp_path / "ten.gro"
    names = ["OW1", "HW2", "HW3", "OW1", "HW2", "HW3", "NA", "C1", "C2", "N"]
    rows = ["    1MOL  %5s%5d%8.3f%8.3f%8.3f\n" % (nm, i + 1, *x[0, i]) for i, nm in enumerate(names)]
    gro.write_text("ten atoms\n   10\n" + "".join(rows) + "   3.00000   3.00000   2.50000\n")
    try:
        import mdtraj  # noqa: F401
    except ImportError:
        lt.process_gro_mdtraj(str(gro), str(tmp_path / "ten.xtc"), str(tmp_path / "out_ten_traj"))
        z = np.load(str(tmp_path / "out_ten_traj.npz"))
        assert list(z["typ"]) == names and z["dims"].shape == (4, 3)
        assert np.array_equal(z["coords"], _xtc_expected(x, 1000.0))
        with pytest.raises(ValueError):          # atom count of the topology must match
            gro.write_text("nine atoms\n    9\n" + "".join(rows[:9]) + "   3.00000   3.00000   2.50000\n")
            lt.process_gro_mdtraj(str(gro), str(tmp_path / "ten.xtc"), str(tmp_path / "out_bad_traj"))


def test_native_dcd_and_psf_readers(tmp_path):
    """The CLI's NAMD route (reference main_gromacs.py:95-96: foo.psf + foo.dcd) without mdtraj: load_traj.read_dcd follows
    the CHARMM/NAMD record layout in either byte order, with and without unit-cell records; load_psf takes the atom
    names of the !NATOM section.  (Parity unpinned: no mdtraj / fixture here, see the reader's docstring.)"""
    import load_traj as lt
    rng = np.random.default_rng(8)
    T, na = 4, 7
    x = rng.uniform(-3, 40, size=(T, na, 3)).astype(np.float32)
    box = np.array([[40.0, 41.0 + t, 39.5] for t in range(T)])
    want = (x * np.float32(0.1)) * np.float32(10)            # Angstrom -> mdtraj's float32 nm -> the reference's * 10
    assert np.abs(want - x).max() < 1e-5
    for big in (False, True):
        for cosines in (True, False):
            p = str(tmp_path / "t.dcd")
            lt.write_dcd(p, x, box, big_endian=big, namd_cosines=cosines)
            c, b = lt.read_dcd(p)
            assert c.dtype == np.float32 and b.dtype == np.float32 and np.array_equal(c, want)
            assert np.allclose(b, box, rtol=1e-6)
    lt.write_dcd(str(tmp_path / "nobox.dcd"), x)
    c, b = lt.read_dcd(str(tmp_path / "nobox.dcd"))
    assert np.array_equal(c, want) and not b.any()
    raw = open(str(tmp_path / "t.dcd"), "rb").read()
    for name, blob in (("cut.dcd", raw[:len(raw) - 9]), ("magic.dcd", raw[:4] + b"VELD" + raw[8:]), ("short.dcd", raw[:50])):
        open(str(tmp_path / name), "wb").write(blob)
        with pytest.raises(ValueError):
            lt.read_dcd(str(tmp_path / name))
    names = ["OH2", "H1", "H2", "SOD", "C1", "N", "O"]
    psf = tmp_path / "t.psf"
    psf.write_text("PSF\n\n       1 !NTITLE\n REMARKS test\n\n%8d !NATOM\n" % na +
                   "".join("%8d W    %-4d TIP3 %-4s %-4s %10.6f %13.4f %11d\n" % (i + 1, i // 3 + 1, nm, "OT", -0.834, 15.9994 + i, 0)
                           for i, nm in enumerate(names)) + "\n%8d !NBOND: bonds\n" % 0)
    got, mass = lt.load_psf(str(psf))
    assert got == names and np.allclose(mass, 15.9994 + np.arange(na))
    try:
        import mdtraj  # noqa: F401
    except ImportError:
        lt.process_gro_mdtraj(str(psf), str(tmp_path / "t.dcd"), str(tmp_path / "out_t_traj"))
        z = np.load(str(tmp_path / "out_t_traj.npz"))
        assert list(z["typ"]) == names and list(z["name"]) == names and np.allclose(z["mass"], mass)
        assert np.array_equal(z["coords"], want) and np.allclose(z["dims"], box, rtol=1e-6) and z["dims"].dtype == np.float32
