"""Trajectory ingest with the reference's entry points (reference load_traj.py).

``load_gro`` (:11-20) and ``process_gro_mdtraj`` (:90-111) keep their signatures and the
``out_<name>_traj.npz`` layout (keys dims, coords, name, mass, typ -- typ holds the atom NAMES,
reference load_traj.py:110).  Decoding of binary trajectories still goes through mdtraj when it is
installed (a native XTC/TRR reader is listed as the next row in SURVEY section 8f); .gro files are
parsed here directly.
"""
import numpy as np


def load_gro(gro):
    """Atom names of a .gro file: columns 10-15 of every atom line (reference load_traj.py:11-20)."""
    with open(gro) as handle:
        rows = handle.readlines()
    return [row[10:15].strip() for row in rows[2:-1]]


def read_gro(gro):
    """One-frame .gro reader: (names, coords in Angstrom float32 (Na,3), box lengths in Angstrom float32 (3,)).

    Box lengths are |a|,|b|,|c| of the (possibly triclinic) box line, like mdtraj's unitcell_lengths."""
    with open(gro) as handle:
        rows = handle.readlines()
    natoms = int(rows[1])
    names = [row[10:15].strip() for row in rows[2:2 + natoms]]
    xyz = np.array([[float(row[20:28]), float(row[28:36]), float(row[36:44])] for row in rows[2:2 + natoms]],
                   dtype=np.float32)
    b = [float(v) for v in rows[2 + natoms].split()]
    if len(b) == 3:
        lengths = np.array(b, dtype=np.float64)
    else:   # v1(x) v2(y) v3(z) v1(y) v1(z) v2(x) v2(z) v3(x) v3(y)
        v1 = np.array([b[0], b[3], b[4]]); v2 = np.array([b[5], b[1], b[6]]); v3 = np.array([b[7], b[8], b[2]])
        lengths = np.array([np.linalg.norm(v1), np.linalg.norm(v2), np.linalg.norm(v3)])
    return names, xyz * np.float32(10), (lengths * 10).astype(np.float32)


def save_traj_npz(output_filename, dims, coords, name, mass=None):
    """Write the traj npz exactly as the reference does (load_traj.py:110): typ = atom names."""
    name = np.asarray(name)
    mass = np.zeros(len(name)) if mass is None else np.asarray(mass)
    np.savez_compressed(output_filename, dims=dims, coords=coords, name=name, mass=mass, typ=name)


def process_gro_mdtraj(topology_filename, trajectory_filename, output_filename):
    """Trajectory + topology -> ``output_filename.npz`` (reference load_traj.py:90-111)."""
    print("processing ", trajectory_filename)
    try:
        import mdtraj as md
    except ImportError as exc:
        if trajectory_filename.endswith(".gro"):
            names, xyz, box = read_gro(trajectory_filename)
            print("saving ", output_filename)
            save_traj_npz(output_filename, box[None, :], xyz[None, :, :], names)
            print('done saving')
            return
        raise ImportError("mdtraj is needed to decode %s (only .gro is parsed natively)" % trajectory_filename) from exc
    t = md.load(trajectory_filename, top=topology_filename)
    coords = t.xyz * 10            # nm -> Angstrom, float32
    dims = t.unitcell_lengths * 10
    name = np.array([a.name for a in t.topology.atoms])
    mass = np.array([a.element.mass for a in t.topology.atoms])
    print("saving ", output_filename)
    save_traj_npz(output_filename, dims, coords, name, mass)
    print('done saving')
