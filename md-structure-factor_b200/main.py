#! /usr/bin/env python
"""Python-3 counterpart of the reference's legacy driver main.py (NAMD inputs, hard-coded names:
W11_large.psf / DH5_423_0.dcd, label o2 -> o2_traj.npz / o2_sf.npz).  The reference script is
Python 2 and calls the old five-argument compute_sf; this keeps its constants and cache logic and
calls the current signature with an orthorhombic cell and 1 Angstrom resolution."""
import os.path

import numpy as np

import dens

top_file = "W11_large.psf"
traj_file = "DH5_423_0.dcd"
label = "o2"
tfname = label + "_traj"
sfname = label + "_sf"


def main():
    if not os.path.isfile(sfname + ".npz"):
        if not os.path.isfile(tfname + ".npz"):
            import load_traj as lt
            print("processing trajectory file " + traj_file)
            lt.process_gro_mdtraj(top_file, traj_file, tfname)
            print('done')
        traj = np.load(tfname + ".npz")
        rad = dens.load_radii(os.path.join(os.path.dirname(os.path.abspath(__file__)), "radii.txt"))
        dens.compute_sf(traj['coords'], traj['dims'], traj['typ'], sfname, rad, np.eye(3), 1.0)
    grid = np.load(sfname + ".npz")['kgridplt']
    print("structure factor grid", grid.shape, "written to", sfname + ".npz")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
