"""GPU post-processing of the structure factor.

``cylindrical_average(D, ucell)`` is the heavy loop of the reference's ``plot2d.PLOT_RAD_NEW`` (plot2d.py:571-638): the
(r, z) map of the theta-averaged, trilinearly interpolated S(q) that the CLI's 2-D pattern is drawn from
(``main_gromacs.py:231`` -> ``Plot_Ewald_triclinic`` -> ``PLOT_RAD_NEW``).  The ring geometry is set up on the host with the
reference's expressions; the 400 x Nz ring averages run in one kernel of ``libmdsf.so`` (``mdsf_cylindrical_average``).
There is no CPU fallback: without the library or a GPU the call raises.
"""
import ctypes as C

import numpy as np

import mdsf_native

THETA_BINS_PER_INV_ANG = 20.     # plot2d.py:585
MIN_THETA_BINS = 1               # plot2d.py:586
RBINS = 400                      # plot2d.py:587


def ring_setup(D, ucell, rbins=RBINS):
    """Axes of D, inverse reciprocal matrix, ring radii, z values, theta samples per ring (plot2d.py:571-574, 590-616)."""
    ucell = np.asarray(ucell, dtype=np.float64)
    X = np.ascontiguousarray(D[:, 0, 0, 0], dtype=np.float64)
    Y = np.ascontiguousarray(D[0, :, 0, 1], dtype=np.float64)
    Z = np.ascontiguousarray(D[0, 0, :, 2], dtype=np.float64)
    a1, a2, a3 = ucell[0], ucell[1], ucell[2]
    b1 = (np.cross(a2, a3)) / (np.dot(a1, np.cross(a2, a3)))
    b2 = (np.cross(a3, a1)) / (np.dot(a2, np.cross(a3, a1)))
    b3 = (np.cross(a1, a2)) / (np.dot(a3, np.cross(a1, a2)))
    b_inv = np.linalg.inv(np.vstack((b1, b2, b3)))
    XR = (X[-1] - X[0]) * ucell[0][0]
    YR = (Y[-1] - Y[0]) * ucell[1][1]
    Rmax = min(XR, YR) / 2.0
    Rmax *= 0.95
    rarr = np.linspace(0.0, Rmax, rbins)
    zar = np.linspace(Z[0], Z[-1], Z.shape[0])
    circ = 2. * np.pi * rarr
    ntheta = np.array([max(int(THETA_BINS_PER_INV_ANG * c), MIN_THETA_BINS) for c in circ], dtype=np.int32)
    return X, Y, Z, b_inv, rarr, zar, ntheta


def cylindrical_average(D, ucell, rbins=RBINS, fill=True, normalize=False, device=0):
    """(oa, rarr, zar): oa[r][z] = mean over theta of S(q) interpolated at (r cos t, r sin t, z) . b_inv.

    ``D`` is ``kgridplt`` of the sf npz (axes in channels 0-2, S(q) in channel 3).  ``fill`` replaces rings that leave the grid
    (NaN) by the smallest finite value (plot2d.py:634-635); ``normalize`` divides by the mean (plot2d.py:637-638)."""
    lib = mdsf_native.load()
    X, Y, Z, b_inv, rarr, zar, ntheta = ring_setup(D, ucell, rbins)
    sf = np.ascontiguousarray(D[..., 3], dtype=np.float64)
    binv = np.ascontiguousarray(b_inv, dtype=np.float64)
    oa = np.empty((rarr.shape[0], zar.shape[0]), dtype=np.float64)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    mdsf_native._check(lib.mdsf_cylindrical_average(int(device), dp(sf), sf.shape[0], sf.shape[1], sf.shape[2], dp(X), dp(Y), dp(Z), dp(binv),
                                                    int(rarr.shape[0]), dp(rarr), ntheta.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    int(zar.shape[0]), dp(zar), dp(oa)))
    if fill:
        mn = np.nanmin(oa)
        oa = np.where(np.isnan(oa), mn, oa)
    if normalize:
        oa = oa / np.average(oa)
    return oa, rarr, zar
