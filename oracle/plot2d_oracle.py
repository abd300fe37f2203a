"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the cylindrical average inside plot2d.PLOT_RAD_NEW
(reference plot2d.py:571-638).  Nothing in the product imports this module.

PARITY UNPINNED: plot2d.py cannot be imported here (matplotlib is absent) and PLOT_RAD_NEW returns nothing -- the averaged
array only feeds matplotlib -- so there is no reference output to pin against.  The restatement below keeps the reference's
statements one for one and calls the same third-party routine (scipy.interpolate.RegularGridInterpolator, linear).
"""
import numpy as np
from scipy.interpolate import RegularGridInterpolator

THETA_BINS_PER_INV_ANG = 20.     # plot2d.py:585
MIN_THETA_BINS = 1               # plot2d.py:586
RBINS = 400                      # plot2d.py:587


def ring_setup(D, ucell, rbins=RBINS):
    """Axes, b_inv, radii, z values and theta counts (plot2d.py:571-574, 590-616)."""
    X = D[:, 0, 0, 0]
    Y = D[0, :, 0, 1]
    Z = D[0, 0, :, 2]
    a1, a2, a3 = ucell[0], ucell[1], ucell[2]
    b1 = (np.cross(a2, a3)) / (np.dot(a1, np.cross(a2, a3)))
    b2 = (np.cross(a3, a1)) / (np.dot(a2, np.cross(a3, a1)))
    b3 = (np.cross(a1, a2)) / (np.dot(a3, np.cross(a1, a2)))
    b_inv = np.linalg.inv(np.vstack((b1, b2, b3)))
    ZBINS = Z.shape[0]
    XR = (X[-1] - X[0]) * ucell[0][0]
    YR = (Y[-1] - Y[0]) * ucell[1][1]
    Rmax = min(XR, YR) / 2.0
    Rmax *= 0.95
    rarr, rspace = np.linspace(0.0, Rmax, rbins, retstep=True)
    zar = np.linspace(Z[0], Z[-1], ZBINS)
    circ = 2. * np.pi * rarr
    ntheta = np.array([max(int(THETA_BINS_PER_INV_ANG * c), MIN_THETA_BINS) for c in circ], dtype=np.int32)
    return X, Y, Z, b_inv, rarr, zar, ntheta


def cylindrical_average(D, ucell, rbins=RBINS, fill=True):
    """oa[r][z] of plot2d.py:604-638 (before the optional division by its mean); returns (oa, rarr, zar)."""
    X, Y, Z, b_inv, rarr, zar, ntheta = ring_setup(D, ucell, rbins)
    SF = D[..., 3]
    ES = RegularGridInterpolator((X, Y, Z), SF, bounds_error=False)
    oa = np.zeros((rarr.shape[0], zar.shape[0]))
    for ir in range(rarr.shape[0]):
        thetas = np.linspace(0.0, np.pi * 2.0, int(ntheta[ir]), endpoint=False)
        t, r, z = np.meshgrid(thetas, rarr[ir], zar)
        xar = r * np.cos(t)
        yar = r * np.sin(t)
        pts = np.vstack((xar.ravel(), yar.ravel(), z.ravel())).T
        MCpts = np.matmul(pts, b_inv)
        oa[ir, :] = np.average(ES(MCpts).reshape(r.shape), axis=1)
    if fill:
        mn = np.nanmin(oa)
        oa = np.where(np.isnan(oa), mn, oa)
    return oa, rarr, zar
