#!/usr/bin/env python
"""Development aid (no GPU needed): a line-by-line Python mirror of K2's pair enumeration (csrc/mdsf_prep.cuh
axis_slot / pairs_of_tile, walked by walk_pairs_warp) and of the way the splat kernel consumes a pair record (csrc/mdsf_splat.cuh), run on the golden cases
and compared with the reference density d1 stored in tests/golden/*.npz.  Validates the index logic (fold images,
corner rule, tile / slab clipping, table offsets) before a GPU run.  Not part of the product or of the test suite.

usage: python tools/emulate_splat.py [case ...] [--lcol 2..5]"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dens_oracle as orc          # noqa: E402
from tests.helpers import CASES, load_case     # noqa: E402


def stamp_segment(ir, A, N, s):
    p0, p1 = ir - A, ir + A
    if s < 0:
        return p0, min(p1, 0)
    if s == 0:
        return max(p0, 0), min(p1, N)
    return max(p0, N), p1


def fold_shift_z(sx, sy, sz, nz, nb, fold_mode):
    if sz == 0:
        return 0
    corner = sx != 0 and sy != 0 and fold_mode == 0
    if sz < 0:
        return nb if (corner and sy != -1) else nz
    return -nb if (corner and sy != 1) else -nz


def emulate_frame(frame, typ, rad, widths, n, dr, nb, ucell, lcol, fold_mode=0):
    nx, ny, nz = (int(v) for v in n)
    ltx, lty = (lcol + 1) >> 1, lcol >> 1
    TX, TY = 1 << ltx, 1 << lty
    lzw = 8 - lcol
    ZW = 1 << lzw
    u = np.asarray(ucell, float)
    cxx = u[0, 0] ** 2 + u[0, 1] ** 2
    cyy = u[1, 0] ** 2 + u[1, 1] ** 2
    gxy = u[0, 0] * u[1, 0] + u[0, 1] * u[1, 1]
    czz = u[2, 2] ** 2
    d = np.zeros((nx, ny, nz))
    npairs = 0
    for a in range(frame.shape[0]):
        r = frame[a].astype(np.float64)
        ir = (r / dr).astype(int)
        A = widths[typ[a]].astype(int)
        Ax, Ay, Az = (int(v) for v in A)
        nel, sig = rad[typ[a]]
        it2 = 1.0 / (2 * sig ** 2)
        amp = nel / sig ** 3
        bx0 = r[0] - (ir[0] - Ax) * dr[0]
        by0 = r[1] - (ir[1] - Ay) * dr[1]
        padx, pady = TX - 1, TY - 1
        T = [0.0] * padx
        for i in range(2 * Ax):
            b = r[0] - (ir[0] - Ax + i) * dr[0]
            T.append(math.exp(-(cxx * b * b + 2.0 * gxy * b * by0) * it2))
        T += [0.0] * (padx + pady)
        for j in range(2 * Ay):
            b = r[1] - (ir[1] - Ay + j) * dr[1]
            T.append(math.exp(-(cyy * b * b - 2.0 * gxy * (j * dr[1]) * bx0) * it2))
        T += [0.0] * pady
        for k in range(2 * Az):
            b = r[2] - (ir[2] - Az + k) * dr[2]
            T.append(amp * math.exp(-(czz * b * b) * it2))
        T = np.array(T)
        cpad = 8 * 2 * Ay + 8
        ctab = np.concatenate([np.ones(cpad), np.array([[math.exp(-(2.0 * gxy * dr[0] * dr[1] * i * j) * it2) for j in range(2 * Ay)]
                                                         for i in range(2 * Ax)]).reshape(-1), np.ones(cpad)])
        tbase = 0
        ex0 = tbase + padx
        ey0 = tbase + 2 * padx + 2 * Ax + pady
        ez0 = tbase + 2 * padx + 2 * Ax + 2 * pady + 2 * Ay
        for sx in (-1, 0, 1):
            xlo, xhi = stamp_segment(ir[0], Ax, nx, sx)
            if xhi <= xlo:
                continue
            dx0, dx1 = xlo - sx * nx, xhi - sx * nx
            ix0 = xlo - (ir[0] - Ax)
            for tX in range(dx0 >> ltx, ((dx1 - 1) >> ltx) + 1):
                X0 = tX << ltx
                cx0 = max(dx0 - X0, 0)
                i0 = ix0 + (X0 + cx0 - dx0)
                for sy in (-1, 0, 1):
                    ylo, yhi = stamp_segment(ir[1], Ay, ny, sy)
                    if yhi <= ylo:
                        continue
                    dy0, dy1 = ylo - sy * ny, yhi - sy * ny
                    jy0 = ylo - (ir[1] - Ay)
                    for tY in range(dy0 >> lty, ((dy1 - 1) >> lty) + 1):
                        Y0 = tY << lty
                        cy0 = max(dy0 - Y0, 0)
                        j0 = jy0 + (Y0 + cy0 - dy0)
                        for sz in (-1, 0, 1):
                            zlo, zhi = stamp_segment(ir[2], Az, nz, sz)
                            if zhi <= zlo:
                                continue
                            shz = fold_shift_z(sx, sy, sz, nz, nb, fold_mode)
                            dz0, dz1 = zlo + shz, zhi + shz
                            kz0 = zlo - (ir[2] - Az)
                            for s in range(dz0 >> lzw, ((dz1 - 1) >> lzw) + 1):
                                Z0 = s << lzw
                                zoff, zend = max(dz0 - Z0, 0), min(dz1 - Z0, ZW)
                                k0 = kz0 + (Z0 + zoff - dz0)
                                ez_idx = ez0 + k0 - zoff
                                ex_idx = ex0 + i0 - cx0
                                ey_idx = ey0 + j0 - cy0
                                ct_idx = cpad + (i0 - cx0) * 2 * Ay + (j0 - cy0)
                                npairs += 1
                                # ---- what the splat warp does with this record: EVERY tile column, no clip test
                                ex = np.array([T[ex_idx + cx] for cx in range(TX)])
                                ey = np.array([T[ey_idx + cy] for cy in range(TY)])
                                cc = np.array([[ctab[ct_idx + cx * 2 * Ay + cy] for cy in range(TY)] for cx in range(TX)])
                                ez = np.array([T[ez_idx + zz] for zz in range(zoff, zend)])
                                exy = ex[:, None] * ey[None, :] * cc
                                x1, y1 = min(X0 + TX, nx), min(Y0 + TY, ny)        # columns outside the grid are never stored
                                d[X0:x1, Y0:y1, Z0 + zoff:Z0 + zend] += (exy[:, :, None] * ez[None, None, :])[:x1 - X0, :y1 - Y0]
    return d, npairs


def main():
    args = sys.argv[1:]
    lcols = [2, 3, 4, 5]
    if "--lcol" in args:
        i = args.index("--lcol")
        lcols = [int(args[i + 1])]
        del args[i:i + 2]
    for name in (args or CASES):
        c = load_case(name)
        r = c["coords"].copy()
        box = orc.rescale_frames(r, c["dims"])
        n, dr = orc.grid_shape(box, c["sres"])
        orc.wrap_frames(r, box)
        widths = orc.half_widths(c["rad"], dr, set(c["typ"]))
        nb = orc.border_cells(widths)
        for lcol in lcols:
            worst = 0.0
            for t in range(r.shape[0]):
                d, npairs = emulate_frame(r[t], c["typ"], c["rad"], widths, n, dr, nb, c["ucell"], lcol)
                worst = max(worst, float(np.abs(d - c["d1"][t]).max() / np.abs(c["d1"][t]).max()))
            print("%-18s grid %s B=%d lcol=%d pairs/frame=%d  max|d-d1|/max = %.2e %s" % (name, tuple(int(v) for v in n), nb, lcol, npairs, worst, "OK" if worst < 1e-13 else "FAIL"))


if __name__ == "__main__":
    main()
