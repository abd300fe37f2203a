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


class NpzFrameStream:
    """Frames of a trajectory npz (reference layout, load_traj.py:110) WITHOUT loading the whole ``coords`` array.

    The reference does ``traj = np.load(...); T = traj['coords']`` (main_gromacs.py:200-202): the complete
    trajectory is inflated into pageable memory before the first frame is used.  Here the ``coords.npy`` zip
    member is inflated incrementally, straight into the caller's (pinned) chunk buffers, so that decoding chunk
    i+1 overlaps the GPU work on chunk i (dens.compute_sf_stream).

    ``shape`` / ``dtype`` describe the full array; ``read_into(buf)`` fills ``buf[:k]`` with the next k frames and
    returns k (0 at the end); ``skip(n)`` drops n frames."""

    def __init__(self, path, key="coords"):
        import zipfile
        self._zip = zipfile.ZipFile(path)
        self._fh = self._zip.open(key + ".npy")
        fmt = np.lib.format
        version = fmt.read_magic(self._fh)
        if version == (1, 0):
            shape, fortran, dtype = fmt.read_array_header_1_0(self._fh)
        else:
            shape, fortran, dtype = fmt.read_array_header_2_0(self._fh)
        if fortran or len(shape) != 3 or shape[2] != 3:
            raise ValueError("%s: coords must be a C-ordered (T, Na, 3) array, got %s" % (path, (shape,)))
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self._frame_bytes = shape[1] * 3 * self.dtype.itemsize
        self.position = 0

    def read_into(self, buf):
        if buf.dtype != self.dtype or buf.shape[1:] != self.shape[1:] or not buf.flags.c_contiguous:
            raise ValueError("chunk buffer must be C-contiguous %s of shape (k, %d, 3)" % (self.dtype, self.shape[1]))
        want = min(buf.shape[0], self.shape[0] - self.position)
        if want <= 0:
            return 0
        view = memoryview(buf[:want].reshape(-1).view(np.uint8))
        got = 0
        while got < len(view):
            n = self._fh.readinto(view[got:])
            if not n:
                raise EOFError("trajectory npz ends inside frame %d" % (self.position + got // max(self._frame_bytes, 1)))
            got += n
        self.position += want
        return want

    def skip(self, nframes):
        nframes = min(int(nframes), self.shape[0] - self.position)
        left = nframes * self._frame_bytes
        while left > 0:
            chunk = self._fh.read(min(left, 16 << 20))
            if not chunk:
                raise EOFError("trajectory npz ends early")
            left -= len(chunk)
        self.position += max(nframes, 0)

    def close(self):
        self._fh.close()
        self._zip.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
