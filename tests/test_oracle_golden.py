"""Pin the CPU oracle (oracle/dens_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py in the build container)."""
import numpy as np
import pytest

from oracle import dens_oracle as orc
from tests.helpers import CASES, GOLDEN, load_case, sf_errors


@pytest.mark.parametrize("name", CASES)
def test_full_path_matches_reference(name):
    c = load_case(name)
    r = c["coords"].copy()
    taps = {}
    out = orc.structure_factor(r, c["dims"].copy(), c["typ"], c["rad"], c["ucell"], c["sres"], taps=taps)
    # integer / exactly-rounded stages: bit-exact
    assert np.array_equal(out["N"], c["ref_N"])
    assert out["L"].dtype == c["ref_L"].dtype and np.array_equal(out["L"], c["ref_L"])
    assert r.dtype == c["coords_after"].dtype and np.array_equal(r, c["coords_after"])
    # density: same arithmetic, summation order of the fold differs -> 1e-13 of the peak
    d1 = np.stack(taps["d1"])
    assert d1.shape == c["d1"].shape
    assert np.abs(d1 - c["d1"]).max() <= 1e-13 * np.abs(c["d1"]).max()
    # S(q): tolerance stated by north_star: 1e-5 per-bin relative; normalised 1e-12
    rel, norm = sf_errors(out["sf"], c["ref_sf"])
    assert rel <= 1e-5 and norm <= 1e-12
    rel, norm = sf_errors(out["sfplt"], c["ref_sfplt"])
    assert rel <= 1e-5 and norm <= 1e-12
    assert np.array_equal(out["kgrid"], c["ref_kgrid"])
    assert np.array_equal(out["kgridplt"][..., :3], c["ref_kgridplt"][..., :3])
    assert out["kgridplt"].shape == c["ref_kgridplt"].shape


def test_fold_matches_reference_including_corner_rule():
    z = np.load(GOLDEN + "/helpers.npz")
    i = 0
    while "fold%d_d0" % i in z.files:
        nb = z["fold%d_nb" % i]
        got = orc.fold_padded(z["fold%d_d0" % i], nb[:3], int(nb[3]), "reference")
        assert np.abs(got - z["fold%d_d1" % i]).max() < 1e-13
        # the periodic fold conserves mass too but differs in the corners
        per = orc.fold_padded(z["fold%d_d0" % i], nb[:3], int(nb[3]), "periodic")
        assert abs(per.sum() - got.sum()) < 1e-9 * got.sum()
        assert np.abs(per - got).max() > 1e-3
        i += 1
    assert i >= 4


def test_centred_view_matches_reference():
    z = np.load(GOLDEN + "/helpers.npz")
    i = 0
    while "dplot%d_in" % i in z.files:
        got = orc.centred_view(z["dplot%d_in" % i])
        assert np.array_equal(got, z["dplot%d_out" % i])
        i += 1
    assert i >= 3


def test_wrap_rule_is_single_shift():
    box = np.array([10.0, 10.0, 10.0])
    r = np.array([[[0.0, 10.0, -10.0], [25.0, -3.0, 12.0]]])
    orc.wrap_frames(r, box)
    assert np.array_equal(r, [[[10.0, 10.0, 0.0], [15.0, 7.0, 2.0]]])


def test_large_system_wrap_quirk_range():
    assert orc.wrapped_atom_range(5, 999999) == (0, 999999)
    assert orc.wrapped_atom_range(5, 1000000) == (0, 5)
    assert orc.wrapped_atom_range(10000, 4000000) == (0, 10000)


def test_normalisation_identity():
    # comment at dens.py:309: dr^3 * sum(d0) / (2 pi)^1.5 == number of electrons
    rad = {"C": (6.0, 0.70)}
    dr = np.array([0.25, 0.25, 0.25])
    w = orc.half_widths(rad, dr, ["C"])
    nb = orc.border_cells(w)
    n = np.array([40, 40, 40])
    d0 = orc.stamp_padded(np.array([[5.1, 4.9, 5.0]]), ["C"], rad, w, n, dr, nb, np.eye(3))
    assert abs(dr.prod() * d0.sum() / (2 * np.pi) ** 1.5 - 6.0) < 1e-6
