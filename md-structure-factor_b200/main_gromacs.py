#! /usr/bin/env python
"""Command-line driver with the reference's flags and artefacts (reference main_gromacs.py).

    python main_gromacs.py -i foo            # foo.gro + foo.trr
    python main_gromacs.py -top foo.gro -traj foo.xtc
    python main_gromacs.py -RC 1000 -RT 4    # random gas      (writes RAND.npz, RND.npz)
    python main_gromacs.py -LX 4 -LY 4 -LZ 4 # simple lattice  (writes lattice_4_4_4.npz)

Artefacts are the reference's: ``out_<name>_traj.npz`` (trajectory cache) and ``out_<name>_sf.npz``
(sf, sfplt, L, N, kgrid, kgridplt); the -fr cache logic is the same (reference main_gromacs.py:190-191).
The structure factor itself is computed by dens.compute_sf on the GPU.  Plotting (plot2d.py of the
reference: Ewald-sphere corrected cross sections, radial averages) is host-side, untimed and not
part of this package; if a ``plot2d`` module is importable it is called exactly like the reference
does, otherwise the run stops after writing the sf npz.
"""
import argparse
import math
import os.path
import platform
import warnings

import numpy as np

import dens

XRAY_WAVELENGTH = 1.54
TRANSFORM_MONOCLINIC = True
STREAM_TRAJECTORY = True     # inflate the traj npz chunk by chunk into pinned buffers, overlapped with the GPU (False: np.load it whole)


def initialize():
    p = argparse.ArgumentParser(description='Calculate 3d Structure Factor')
    p.add_argument('-i', '--input', default='', type=str, help='basename of topology and trajectory')
    p.add_argument('-top', '--topology', default='', type=str, help='topology file')
    p.add_argument('-traj', '--trajectory', default='', type=str, help='trajectory file')
    p.add_argument('--cscale', default=1, type=float, help='colour-map scale of the plots')
    p.add_argument('--lcscale', default=1, type=float, help='colour-map scale of the log plots')
    p.add_argument('-fi', '--first_frame', default=0, type=int, help='first frame')
    p.add_argument('-e', '--end_frame', default=-1, type=int, help='end frame (python slice end; -1 drops the last frame)')
    p.add_argument('-tf', '--traj_format', default='gromacs', type=str, help='trajectory format: gromacs or namd')
    p.add_argument('-fr', '--force_recompute', default=0, type=int, help='>=1: recompute SF; >=2: also re-read the trajectory')
    p.add_argument('-RC', '--random_counts', default=0, type=int, help='number of random particles (random-gas mode)')
    p.add_argument('-RT', '--random_timesteps', default=1, type=int, help='number of random frames')
    p.add_argument('-RL', '--random_label', default='R3', type=str, help='label of the random particles')
    p.add_argument('-LX', '--lattice_x', default=0, type=int, help='lattice points along x (lattice mode)')
    p.add_argument('-LY', '--lattice_y', default=1, type=int, help='lattice points along y')
    p.add_argument('-LZ', '--lattice_z', default=1, type=int, help='lattice points along z')
    p.add_argument('-LL', '--lattice_label', default='R3', type=str, help='label of the lattice particles')
    p.add_argument('-SR', '--spatial_resolution', default=1.0, type=float, help='grid resolution in Angstrom')
    p.add_argument('-RN', '--random_noise', default=0, type=int, help='>0: replace the density by random noise (test mode)')
    p.add_argument('-RS', '--random_seed', default=1, type=int, help='random seed')
    p.add_argument('-NBR', '--number_bins_rad', default=0, type=int, help='number of radial bins for the plots')
    p.add_argument('-ct', '--cell_theta', default=120, type=float, help='cell angle theta in degrees')
    p.add_argument('-nocbar', '--nocolorbar', action='store_true', help='plots without colour bar')
    p.add_argument('-scale_factor', default=3.1, type=float, help='maximum colour-bar value of the XRD pattern')
    p.add_argument('-manuscript_format', action='store_true', help='plot formatting of Coscia et al.')
    p.add_argument('--device', default=0, type=int, help='CUDA device ordinal (this implementation only)')
    p.add_argument('--exact-fold', action='store_true', help='use the exact periodic fold instead of the reference\'s corner rule')
    return p


def synthetic_system(args):
    """Lattice (-LX/-LY/-LZ) and random-gas (-RC) inputs, float64 like the reference builds them
    (reference main_gromacs.py:122-186).  Returns (coords, dims, typ, sfname, cache name or None)."""
    nlat = args.lattice_x * args.lattice_y * args.lattice_z
    box = 100.0
    if nlat > 0:
        dims = np.ones((1, 3)) * box
        axes = [np.linspace(0, box, n, endpoint=False) for n in (args.lattice_x, args.lattice_y, args.lattice_z)]
        gx, gy, gz = np.meshgrid(*axes)
        coords = np.zeros((1, nlat, 3))
        coords[0, :, 0], coords[0, :, 1], coords[0, :, 2] = gx.reshape(nlat), gy.reshape(nlat), gz.reshape(nlat)
        typ = np.zeros(nlat, dtype=object)
        typ[:] = args.lattice_label
        name = 'lattice_%d_%d_%d' % (args.lattice_x, args.lattice_y, args.lattice_z)
        return coords, dims, typ, name, name
    dims = np.ones((args.random_timesteps, 3)) * box
    coords = np.random.random((args.random_timesteps, args.random_counts, 3)) * dims[0, :]
    typ = np.zeros(args.random_counts, dtype=object)
    typ[:] = args.random_label
    return coords, dims, typ, 'RND', 'RAND'


def main(argv=None):
    args = initialize().parse_args(argv)
    location = os.path.realpath(os.path.join(os.getcwd(), os.path.dirname(__file__)))
    warnings.simplefilter("ignore", RuntimeWarning)

    theta = args.cell_theta * math.pi / 180.0
    ucell = np.array([[1, 0, 0], [np.cos(theta), np.sin(theta), 0], [0, 0, 1]])
    np.random.seed(args.random_seed)        # the reference assigns np.random.seed instead of calling it; we seed
    dens.theta = theta
    dens.DEVICE = args.device
    if args.exact_fold:
        dens.FOLD_MODE = "periodic"
    if args.random_noise > 0:
        dens.RANDOM_NOISE = args.random_noise

    ext_top = {"gromacs": ".gro", "namd": ".psf"}
    ext_traj = {"gromacs": ".trr", "namd": ".dcd"}
    if len(args.input) > 0:
        basename = args.input
        top_file = args.input + ext_top[args.traj_format.lower()]
        traj_file = args.input + ext_traj[args.traj_format.lower()]
    else:
        top_file, traj_file = args.topology, args.trajectory
        basename = args.topology.rsplit('.', 1)[0]
    print("running on", platform.system(), platform.release(), platform.version())

    tfname = "out_" + basename + "_traj"
    sfname = "out_" + basename + "_sf"
    rad_file = "%s/radii.txt" % location

    if args.lattice_x * args.lattice_y * args.lattice_z > 0 or args.random_counts > 0:
        coords, dims, typ, sfname, cache = synthetic_system(args)
        print("saving...")
        np.savez_compressed(cache, dims=dims, coords=coords, name=np.zeros(len(typ), dtype=object), typ=typ)
        rad = dens.load_radii(rad_file)
        print("computing SF...")
        dens.compute_sf(coords, dims, typ, sfname, rad, ucell, args.spatial_resolution)
    elif args.force_recompute > 0 or not os.path.isfile(sfname + ".npz"):
        if args.force_recompute > 1 or not os.path.isfile(tfname + ".npz"):
            import load_traj as lt
            print("processing trajectory file " + traj_file)
            lt.process_gro_mdtraj(top_file, traj_file, tfname)
            print('done')
        traj = np.load(tfname + ".npz")
        rad = dens.load_radii(rad_file)
        mono = theta if (TRANSFORM_MONOCLINIC and theta != 90.0) else None      # (radians vs 90: always true, as in the reference)
        if mono is not None:
            print("transforming coordinates to monoclinic cell (theta={0:f} deg)".format(theta * 180.0 / np.pi))
        if STREAM_TRAJECTORY:
            # same result as the in-memory route below, but the coords member is inflated chunk by chunk into pinned
            # buffers while the GPU works on the previous chunk (load_traj.NpzFrameStream, dens.compute_sf_stream)
            import load_traj as lt
            with lt.NpzFrameStream(tfname + ".npz") as frames:
                dens.compute_sf_stream(frames, traj['dims'], traj['typ'], sfname, rad, ucell, args.spatial_resolution,
                                       first_frame=args.first_frame, end_frame=args.end_frame, monoclinic_theta=mono)
        else:
            T = traj['coords']
            if mono is not None:
                T[..., 1] = T[..., 1] / np.sin(theta)
                T[..., 0] = T[..., 0] - T[..., 1] * np.cos(theta)
            dens.compute_sf(T[args.first_frame:args.end_frame, ...], traj['dims'][args.first_frame:args.end_frame, ...],
                            traj['typ'], sfname, rad, ucell, args.spatial_resolution)

    print("reloading SF...")
    grid = np.load(sfname + ".npz")['kgridplt']
    try:
        import plot2d as p2d
    except ImportError:
        print("plot2d is not installed: wrote %s.npz (kgridplt %s); plotting skipped" % (sfname, grid.shape))
        return 0
    p2d.NBINSRAD = args.number_bins_rad
    p2d.theta = theta
    if args.nocolorbar:
        p2d.colorbar = False
    p2d.mainlabel = basename
    p2d.path = basename + "_plots/Ewald_Corrected/"
    print("making plots...")
    p2d.Plot_Ewald_triclinic(grid, XRAY_WAVELENGTH, ucell, factor=args.scale_factor, format=args.manuscript_format)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
