"""Trajectory ingest with the reference's entry points (reference load_traj.py).

``load_gro`` (:11-20) and ``process_gro_mdtraj`` (:90-111) keep their signatures and the
``out_<name>_traj.npz`` layout (keys dims, coords, name, mass, typ -- typ holds the atom NAMES,
reference load_traj.py:110).  With mdtraj installed every format goes through it, as in the reference;
without it, .gro (one or many frames), .trr, .xtc and NAMD .psf/.dcd are decoded here directly (the xtc3
coordinate blocks by libmdsf_io on all host cores).
"""
import ctypes
import os

import numpy as np

import npz_writer

TRAJ_PIECE = 1 << 20      # raw bytes per independently deflated piece of a traj npz member (about one c2 frame)
IO_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmdsf_io.so")
IO_EXPORTS = ["mdsf_io_inflate_pieces", "mdsf_io_xtc_decode_frames", "mdsf_io_abi_version"]
_io = None


def _io_lib():
    """libmdsf_io.so (include/mdsf_io.h): built by csrc/build.sh next to this file."""
    global _io
    if _io is None:
        if not os.path.exists(IO_LIB_PATH):
            raise ImportError("%s is missing: run md-structure-factor_b200/csrc/build.sh" % IO_LIB_PATH)
        lib = ctypes.CDLL(IO_LIB_PATH)
        p64 = ctypes.POINTER(ctypes.c_int64)
        lib.mdsf_io_inflate_pieces.restype = ctypes.c_int
        lib.mdsf_io_inflate_pieces.argtypes = [ctypes.c_int, ctypes.c_int64, p64, p64, ctypes.POINTER(ctypes.c_void_p), p64, ctypes.c_int]
        lib.mdsf_io_abi_version.restype = ctypes.c_int
        lib.mdsf_io_xtc_decode_frames.restype = ctypes.c_int
        lib.mdsf_io_xtc_decode_frames.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, p64, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_int]
        _io = lib
    return _io


def load_gro(gro):
    """Atom names of a .gro file: columns 10-15 of every atom line (reference load_traj.py:11-20)."""
    with open(gro) as handle:
        rows = handle.readlines()
    return [row[10:15].strip() for row in rows[2:-1]]


def _gro_width(row):
    """Width of the coordinate fields of a .gro atom line: GROMACS writes %(n+5).(n)f (8 columns at the default 3
    decimals, 10 with -ndec 5, ...); like GROMACS and mdtraj, take it from the spacing of the decimal points."""
    p1 = row.find(".", 20)
    p2 = row.find(".", p1 + 1) if p1 >= 0 else -1
    return p2 - p1 if p1 >= 0 and p2 > p1 else 8


def _gro_xyz(rows):
    w = _gro_width(rows[0]) if rows else 8
    return np.array([[float(row[20:20 + w]), float(row[20 + w:20 + 2 * w]), float(row[20 + 2 * w:20 + 3 * w])] for row in rows],
                    dtype=np.float32)


def read_gro(gro):
    """One-frame .gro reader: (names, coords in Angstrom float32 (Na,3), box lengths in Angstrom float32 (3,)).

    Box lengths are |a|,|b|,|c| of the (possibly triclinic) box line, like mdtraj's unitcell_lengths."""
    with open(gro) as handle:
        rows = handle.readlines()
    natoms = int(rows[1])
    names = [row[10:15].strip() for row in rows[2:2 + natoms]]
    xyz = _gro_xyz(rows[2:2 + natoms])
    b = [float(v) for v in rows[2 + natoms].split()]
    if len(b) == 3:
        lengths = np.array(b, dtype=np.float64)
    else:   # v1(x) v2(y) v3(z) v1(y) v1(z) v2(x) v2(z) v3(x) v3(y)
        v1 = np.array([b[0], b[3], b[4]]); v2 = np.array([b[5], b[1], b[6]]); v3 = np.array([b[7], b[8], b[2]])
        lengths = np.array([np.linalg.norm(v1), np.linalg.norm(v2), np.linalg.norm(v3)])
    return names, xyz * np.float32(10), (lengths * 10).astype(np.float32)


def read_gro_frames(gro):
    """Multi-frame .gro reader (GROMACS writes trajectories as concatenated .gro frames): (names, coords in Angstrom
    float32 (T, Na, 3), box lengths in Angstrom float32 (T, 3)).  A one-frame file gives T = 1."""
    with open(gro) as handle:
        rows = handle.readlines()
    frames, boxes, names = [], [], None
    pos = 0
    while pos + 2 < len(rows) and rows[pos + 1].strip():
        natoms = int(rows[pos + 1])
        body = rows[pos + 2:pos + 2 + natoms]
        if names is None:
            names = [row[10:15].strip() for row in body]
        elif len(body) != len(names):
            raise ValueError("%s: frame %d has %d atoms, frame 0 has %d" % (gro, len(frames), len(body), len(names)))
        frames.append(_gro_xyz(body))
        b = [float(v) for v in rows[pos + 2 + natoms].split()]
        if len(b) == 3:
            boxes.append(np.array(b, dtype=np.float64))
        else:   # v1(x) v2(y) v3(z) v1(y) v1(z) v2(x) v2(z) v3(x) v3(y)
            v1 = np.array([b[0], b[3], b[4]]); v2 = np.array([b[5], b[1], b[6]]); v3 = np.array([b[7], b[8], b[2]])
            boxes.append(np.array([np.linalg.norm(v1), np.linalg.norm(v2), np.linalg.norm(v3)]))
        pos += natoms + 3
    if not frames:
        raise ValueError("%s holds no .gro frame" % gro)
    return names, np.stack(frames) * np.float32(10), (np.stack(boxes) * 10).astype(np.float32)


TRR_MAGIC = 1993
TRR_VERSION = b"GMX_trn_file"


def read_trr(path):
    """GROMACS .trr reader (XDR, big-endian; single or double precision): (coords in Angstrom float32 (T, Na, 3),
    box lengths in Angstrom float32 (T, 3), times in ps).  Frames without coordinates (velocity/force-only output
    steps) are skipped.  Layout per frame, as written by GROMACS' trnio: magic 1993, the version string
    "GMX_trn_file" (as int 13, int 12, 12 bytes), the ten block sizes ir, e, box, vir, pres, top, sym, x, v, f, then
    natoms, step, nre, t, lambda (reals), then the blocks box(3x3), vir, pres, x, v, f of `real`s in nm.
    PARITY UNPINNED: the reference decodes .trr through mdtraj (load_traj.py:94), which is not installed here, and ships
    no .trr fixture; this reader is checked against a writer of the same published layout (tests/test_host_logic.py)."""
    data = np.fromfile(path, dtype=np.uint8)
    pos, n = 0, data.size
    xyz, boxes, times = [], [], []
    be32 = np.dtype(">i4")
    while pos < n:
        if pos + 8 + 4 + len(TRR_VERSION) > n:
            raise ValueError("%s: truncated frame header at byte %d" % (path, pos))
        magic, slen = np.frombuffer(data, be32, 2, pos)
        if magic != TRR_MAGIC or slen != len(TRR_VERSION) + 1:
            raise ValueError("%s: not a .trr frame at byte %d (magic %d)" % (path, pos, magic))
        pos += 8
        strlen = int(np.frombuffer(data, be32, 1, pos)[0])
        pos += 4
        if bytes(data[pos:pos + strlen]) != TRR_VERSION:
            raise ValueError("%s: unexpected version string at byte %d" % (path, pos))
        pos += (strlen + 3) // 4 * 4
        ir, e, box, vir, pres, top, sym, xs, vs, fs, natoms, step, nre = (int(v) for v in np.frombuffer(data, be32, 13, pos))
        pos += 52
        if box:
            real = box // 9
        elif natoms and (xs or vs or fs):
            real = (xs or vs or fs) // (3 * natoms)
        else:
            raise ValueError("%s: frame at step %d has neither box nor vectors" % (path, step))
        if real not in (4, 8):
            raise ValueError("%s: real size %d is neither float nor double" % (path, real))
        rt = np.dtype(">f%d" % real)
        t = float(np.frombuffer(data, rt, 1, pos)[0])
        pos += 2 * real                                   # t, lambda
        pos += ir + e
        bvec = np.frombuffer(data, rt, 9, pos).reshape(3, 3).astype(np.float64) if box else np.zeros((3, 3))
        pos += box + vir + pres + top + sym
        if xs:
            if xs != natoms * 3 * real:
                raise ValueError("%s: coordinate block of %d bytes for %d atoms" % (path, xs, natoms))
            xyz.append(np.frombuffer(data, rt, natoms * 3, pos).reshape(natoms, 3).astype(np.float32))
            boxes.append(np.sqrt((bvec * bvec).sum(axis=1)))
            times.append(t)
        pos += xs + vs + fs
        if pos > n:
            raise ValueError("%s: truncated frame at step %d" % (path, step))
    if not xyz:
        raise ValueError("%s holds no frame with coordinates" % path)
    return np.stack(xyz) * np.float32(10), (np.stack(boxes) * 10).astype(np.float32), np.array(times)


def write_trr(path, coords_nm, box_nm, times=None, double=False, velocities=None):
    """Writer of the same layout (tests, synthetic trajectories): coords (T, Na, 3) and box vectors (T, 3, 3) or lengths
    (T, 3) in nm."""
    coords_nm = np.asarray(coords_nm)
    T, natoms = coords_nm.shape[:2]
    box_nm = np.asarray(box_nm, dtype=np.float64)
    if box_nm.ndim == 2:
        box_nm = np.stack([np.diag(b) for b in box_nm])
    real = 8 if double else 4
    rt = np.dtype(">f%d" % real)
    with open(path, "wb") as fh:
        for it in range(T):
            vsize = natoms * 3 * real if velocities is not None else 0
            head = np.array([TRR_MAGIC, len(TRR_VERSION) + 1, len(TRR_VERSION)], dtype=">i4").tobytes() + TRR_VERSION
            sizes = np.array([0, 0, 9 * real, 0, 0, 0, 0, natoms * 3 * real, vsize, 0, natoms, it, 0], dtype=">i4").tobytes()
            tl = np.array([0.0 if times is None else times[it], 0.0], dtype=rt).tobytes()
            fh.write(head + sizes + tl + box_nm[it].astype(rt).tobytes() + coords_nm[it].astype(rt).tobytes())
            if velocities is not None:
                fh.write(np.asarray(velocities[it]).astype(rt).tobytes())


def load_psf(psf):
    """(atom names, masses) from the !NATOM section of a CHARMM/NAMD/X-PLOR .psf topology: one line per atom,
    ``id segment resid resname NAME type charge mass ...`` (whitespace separated, standard or EXT widths).  The NAMD
    route of the reference CLI (main_gromacs.py:95-96: foo.psf + foo.dcd) reads it through mdtraj."""
    with open(psf) as handle:
        rows = handle.readlines()
    for i, row in enumerate(rows):
        if "!NATOM" in row:
            natoms = int(row.split()[0])
            body = rows[i + 1:i + 1 + natoms]
            if len(body) != natoms:
                raise ValueError("%s: !NATOM announces %d atoms, %d lines follow" % (psf, natoms, len(body)))
            cols = [r.split() for r in body]
            if any(len(c) < 8 for c in cols):
                raise ValueError("%s: atom line with fewer than 8 fields" % psf)
            return [c[4] for c in cols], np.array([float(c[7]) for c in cols])
    raise ValueError("%s holds no !NATOM section" % psf)


def read_dcd(path):
    """CHARMM/NAMD .dcd reader: (coords in Angstrom float32 (T, Na, 3), box lengths in Angstrom float32 (T, 3)).

    Fortran unformatted records with 4-byte markers, either endianness: an 84-byte header ``CORD`` + 20 control words
    (frames, first step, step stride, ..., fixed atoms [8], time step [9], unit-cell flag [10], 4-D flag [11], ...,
    CHARMM version [19]), the title record, the atom count, then per frame an optional unit-cell record of six
    doubles (A, gamma, B, beta, alpha, C: lengths at 0, 2, 5) and the X, Y, Z records of natoms float32 each (and a
    fourth one when the 4-D flag is set).  The file is in Angstrom; mdtraj hands the reference nanometres in float32
    and the reference multiplies by ten again (load_traj.py:97-98), so the same two float32 roundings are applied.
    Files with fixed atoms (later frames hold the free atoms only) are not supported.
    PARITY UNPINNED: no mdtraj and no .dcd fixture here; checked against `write_dcd` (tests/test_host_logic.py)."""
    data = np.fromfile(path, dtype=np.uint8)
    if data.size < 92:
        raise ValueError("%s: too short for a .dcd header" % path)
    for order in ("<", ">"):
        if int(np.frombuffer(data, order + "i4", 1, 0)[0]) == 84:
            break
    else:
        raise ValueError("%s: no 84-byte header record (64-bit record markers are not supported)" % path)
    i4, f4, f8 = np.dtype(order + "i4"), np.dtype(order + "f4"), np.dtype(order + "f8")
    if bytes(data[4:8]) != b"CORD":
        raise ValueError("%s: header does not start with CORD" % path)
    ctl = np.frombuffer(data, i4, 20, 8)
    if int(ctl[8]) != 0:
        raise NotImplementedError("%s: %d fixed atoms (partial frames) are not supported" % (path, int(ctl[8])))
    has_cell, has_4d = int(ctl[10]) != 0, int(ctl[11]) != 0 and int(ctl[19]) != 0
    pos = 92

    def record(at):
        if at + 4 > data.size:
            raise ValueError("%s: truncated record at byte %d" % (path, at))
        n = int(np.frombuffer(data, i4, 1, at)[0])
        if n < 0 or at + 8 + n > data.size or int(np.frombuffer(data, i4, 1, at + 4 + n)[0]) != n:
            raise ValueError("%s: broken record at byte %d" % (path, at))
        return at + 4, n

    at, n = record(pos)                                # title
    pos = at + n + 4
    at, n = record(pos)                                # atom count
    if n != 4:
        raise ValueError("%s: atom-count record of %d bytes" % (path, n))
    natoms = int(np.frombuffer(data, i4, 1, at)[0])
    pos = at + n + 4
    xyz, boxes = [], []
    while pos < data.size:
        cell = np.zeros(3)
        if has_cell:
            at, n = record(pos)
            if n != 48:
                raise ValueError("%s: unit-cell record of %d bytes" % (path, n))
            c = np.frombuffer(data, f8, 6, at)
            cell = np.array([c[0], c[2], c[5]])
            pos = at + n + 4
        frame = np.empty((natoms, 3), dtype=np.float32)
        for d in range(3 + (1 if has_4d else 0)):
            at, n = record(pos)
            if n != 4 * natoms:
                raise ValueError("%s: coordinate record of %d bytes for %d atoms" % (path, n, natoms))
            if d < 3:
                frame[:, d] = np.frombuffer(data, f4, natoms, at)
            pos = at + n + 4
        xyz.append(frame)
        boxes.append(cell)
    if not xyz:
        raise ValueError("%s holds no frame" % path)
    nm, ten = np.float32(0.1), np.float32(10)
    return (np.stack(xyz) * nm) * ten, (np.stack(boxes).astype(np.float32) * nm) * ten


def write_dcd(path, coords_A, box_A=None, big_endian=False, namd_cosines=True):
    """.dcd writer of the same layout (tests, synthetic trajectories): coords (T, Na, 3) and box lengths (T, 3) in Angstrom
    (orthorhombic: the three angle slots hold cos 90 = 0 as NAMD writes them, or 90 degrees as CHARMM does)."""
    coords_A = np.asarray(coords_A, dtype=np.float32)
    T, natoms = coords_A.shape[:2]
    o = ">" if big_endian else "<"

    def rec(payload):
        n = np.array([len(payload)], dtype=o + "i4").tobytes()
        return n + payload + n

    ctl = np.zeros(20, dtype=o + "i4")
    ctl[0], ctl[1], ctl[2], ctl[3] = T, 0, 1, T
    ctl[10] = 1 if box_A is not None else 0
    ctl[19] = 24
    ctl_bytes = bytearray(ctl.tobytes())
    ctl_bytes[36:40] = np.array([1.0], dtype=o + "f4").tobytes()        # time step
    title = b"written by mdsf-b200".ljust(80)
    with open(path, "wb") as fh:
        fh.write(rec(b"CORD" + bytes(ctl_bytes)))
        fh.write(rec(np.array([1], dtype=o + "i4").tobytes() + title))
        fh.write(rec(np.array([natoms], dtype=o + "i4").tobytes()))
        ang = 0.0 if namd_cosines else 90.0
        for it in range(T):
            if box_A is not None:
                a, b, c = (float(v) for v in box_A[it])
                fh.write(rec(np.array([a, ang, b, ang, ang, c], dtype=o + "f8").tobytes()))
            for d in range(3):
                fh.write(rec(np.ascontiguousarray(coords_A[it, :, d]).astype(o + "f4").tobytes()))


XTC_MAGIC = 1995
XTC_MAGICINTS = [0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406,
                 512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642,
                 26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127,
                 524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491,
                 6658042, 8388607, 10568983, 13316085, 16777216]
XTC_FIRSTIDX = 9


def _scan_xtc(data, path):
    """Walk the frame headers of an .xtc image: (coordinate-block offsets int64 (T,), natoms, box lengths in Angstrom
    float32 (T, 3), times)."""
    n, pos = data.size, 0
    be32, bf32 = np.dtype(">i4"), np.dtype(">f4")
    offs, boxes, times = [], [], []
    natoms = None
    while pos < n:
        if pos + 56 > n:
            raise ValueError("%s: truncated frame header at byte %d" % (path, pos))
        magic, na, step = (int(v) for v in np.frombuffer(data, be32, 3, pos))
        if magic != XTC_MAGIC:
            raise ValueError("%s: not an .xtc frame at byte %d (magic %d)" % (path, pos, magic))
        if na <= 0:
            raise ValueError("%s: frame at step %d declares %d atoms" % (path, step, na))
        if natoms is None:
            natoms = na
        elif na != natoms:
            raise ValueError("%s: frame at step %d has %d atoms, the first one %d" % (path, step, na, natoms))
        times.append(float(np.frombuffer(data, bf32, 1, pos + 12)[0]))
        bvec = np.frombuffer(data, bf32, 9, pos + 16).reshape(3, 3).astype(np.float64)
        boxes.append(np.sqrt((bvec * bvec).sum(axis=1)))
        pos += 52
        offs.append(pos)
        if na <= 9:
            pos += 4 + 12 * na
        else:
            if pos + 40 > n:
                raise ValueError("%s: truncated coordinate header at step %d" % (path, step))
            nbytes = int(np.frombuffer(data, be32, 1, pos + 36)[0])
            if nbytes < 0 or na > 8 * nbytes:          # every atom takes more than one bit of the stream
                raise ValueError("%s: byte count %d cannot hold %d atoms at step %d" % (path, nbytes, na, step))
            pos += 40 + (nbytes + 3) // 4 * 4
        if pos > n:
            raise ValueError("%s: truncated frame at step %d" % (path, step))
    if not offs:
        raise ValueError("%s holds no frame" % path)
    return np.asarray(offs, dtype=np.int64), natoms, (np.stack(boxes) * 10).astype(np.float32), np.array(times)


def _decode_xtc(data, off, natoms, out, threads, path, first):
    """xtc3 blocks at ``off`` -> ``out`` (k, natoms, 3) float32 nm, through libmdsf_io."""
    rc = _io_lib().mdsf_io_xtc_decode_frames(data.ctypes.data, data.size, len(off), off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                             natoms, out.ctypes.data, int(threads or 0))
    if rc != 0:
        raise ValueError("%s: corrupt coordinate block in frame %d" % (path, first - rc - 1))


def read_xtc(path, threads=0):
    """GROMACS .xtc reader: (coords in Angstrom float32 (T, Na, 3), box lengths in Angstrom float32 (T, 3), times ps).

    Frame layout (XDR, big-endian): magic 1995, natoms, step, time f32, box 3x3 f32 (nm), then the coordinate block:
    natoms again, and either natoms*3 plain floats (natoms <= 9) or precision f32, minint[3], maxint[3], smallidx,
    byte count and the xtc3 bit stream padded to four bytes.  The headers are walked here; the coordinate blocks are
    decoded by libmdsf_io (mdsf_io_xtc_decode_frames, include/mdsf_io.h), one frame per host thread.
    PARITY UNPINNED: the reference decodes .xtc through mdtraj (load_traj.py:94), which is not installed here, and
    ships no .xtc fixture; decoder and `write_xtc` are separate restatements of the published scheme checked against
    each other (tests/test_host_logic.py)."""
    data = np.fromfile(path, dtype=np.uint8)
    off, natoms, box, times = _scan_xtc(data, path)
    out = np.empty((len(off), natoms, 3), dtype=np.float32)
    _decode_xtc(data, off, natoms, out, threads, path, 0)
    return out * np.float32(10), box, times


class _BitWriter:
    def __init__(self):
        self.acc, self.nacc, self.out = 0, 0, bytearray()

    def bits(self, nbits, value):                 # most significant bit first
        self.acc = (self.acc << nbits) | (int(value) & ((1 << nbits) - 1))
        self.nacc += nbits
        while self.nacc >= 8:
            self.nacc -= 8
            self.out.append((self.acc >> self.nacc) & 0xff)
        self.acc &= (1 << self.nacc) - 1

    def ints3(self, nbits, sizes, vals):          # ((a * s1) + b) * s2 + c, low byte first
        v = (int(vals[0]) * sizes[1] + int(vals[1])) * sizes[2] + int(vals[2])
        assert 0 <= v < (1 << nbits)
        while nbits > 8:
            self.bits(8, v & 0xff)
            v >>= 8
            nbits -= 8
        self.bits(nbits, v)

    def done(self):
        if self.nacc:
            self.out.append((self.acc << (8 - self.nacc)) & 0xff)
        return bytes(self.out)


def _xtc_pack(ints, precision):
    """xtc3 coordinate block of one frame (ints (Na, 3) = round(x * precision), Na > 9): the encoder side of the
    scheme read_xtc decodes -- bounding-box coded group leaders, runs of up to eight small-offset atoms with the
    leader swapped behind the first of them, small range following the nearest-neighbour distances."""
    M = XTC_MAGICINTS
    size = len(ints)
    c = [list(map(int, row)) for row in ints]
    minint = [min(r[d] for r in c) for d in range(3)]
    maxint = [max(r[d] for r in c) for d in range(3)]
    mindiff = min(sum(abs(c[i][d] - c[i - 1][d]) for d in range(3)) for i in range(1, size))
    sizeint = [maxint[d] - minint[d] + 1 for d in range(3)]
    if max(sizeint) > 0xffffff:
        bitsizeint, bitsize = [s.bit_length() for s in sizeint], 0
    else:
        bitsize = (sizeint[0] * sizeint[1] * sizeint[2]).bit_length()
    smallidx = XTC_FIRSTIDX
    while smallidx < len(M) - 1 and M[smallidx] < mindiff:
        smallidx += 1
    smallidx0 = smallidx
    maxidx = min(len(M) - 1, smallidx + 8)
    minidx = maxidx - 8
    smaller = M[max(XTC_FIRSTIDX, smallidx - 1)] // 2
    smallnum = M[smallidx] // 2
    larger = M[maxidx] // 2
    bw = _BitWriter()
    i, prevrun, prev = 0, -1, [0, 0, 0]
    while i < size:
        is_small = False
        if smallidx < maxidx and i >= 1 and all(abs(c[i][d] - prev[d]) < larger for d in range(3)):
            is_smaller = 1
        elif smallidx > minidx:
            is_smaller = -1
        else:
            is_smaller = 0
        if i + 1 < size and all(abs(c[i][d] - c[i + 1][d]) < smallnum for d in range(3)):
            c[i], c[i + 1] = c[i + 1], c[i]        # water: the oxygen goes behind its first hydrogen
            is_small = True
        lead = [c[i][d] - minint[d] for d in range(3)]
        if bitsize == 0:
            for d in range(3):
                bw.bits(bitsizeint[d], lead[d])
        else:
            bw.ints3(bitsize, sizeint, lead)
        prev = c[i]
        i += 1
        run = []
        if not is_small and is_smaller == -1:
            is_smaller = 0
        while is_small and len(run) < 8:
            if is_smaller == -1 and sum((c[i][d] - prev[d]) ** 2 for d in range(3)) >= smaller * smaller:
                is_smaller = 0
            run.append([c[i][d] - prev[d] + smallnum for d in range(3)])
            prev = c[i]
            i += 1
            is_small = i < size and all(abs(c[i][d] - prev[d]) < smallnum for d in range(3))
        nrun = 3 * len(run)
        if nrun != prevrun or is_smaller != 0:
            prevrun = nrun
            bw.bits(1, 1)
            bw.bits(5, nrun + is_smaller + 1)
        else:
            bw.bits(1, 0)
        for r in run:
            bw.ints3(smallidx, [M[smallidx]] * 3, r)
        if is_smaller != 0:
            smallidx += is_smaller
            if is_smaller < 0:
                smallnum = smaller
                smaller = M[smallidx - 1] // 2
            else:
                smaller = smallnum
                smallnum = M[smallidx] // 2
    payload = bw.done()
    head = np.array([precision], dtype=">f4").tobytes() + np.array(minint + maxint + [smallidx0, len(payload)], dtype=">i4").tobytes()
    return head + payload + b"\0" * (-len(payload) % 4)


def write_xtc(path, coords_nm, box_nm, times=None, precision=1000.0):
    """.xtc writer (tests, synthetic trajectories): coords (T, Na, 3) and box vectors (T, 3, 3) or lengths (T, 3) in nm."""
    coords_nm = np.asarray(coords_nm, dtype=np.float32)
    T, natoms = coords_nm.shape[:2]
    box_nm = np.asarray(box_nm, dtype=np.float64)
    if box_nm.ndim == 2:
        box_nm = np.stack([np.diag(b) for b in box_nm])
    with open(path, "wb") as fh:
        for it in range(T):
            fh.write(np.array([XTC_MAGIC, natoms, it], dtype=">i4").tobytes())
            fh.write(np.array([0.0 if times is None else times[it]], dtype=">f4").tobytes())
            fh.write(box_nm[it].astype(">f4").tobytes())
            fh.write(np.array([natoms], dtype=">i4").tobytes())
            if natoms <= 9:
                fh.write(coords_nm[it].astype(">f4").tobytes())
            else:
                x = coords_nm[it].astype(np.float64) * np.float32(precision)
                ints = np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)
                fh.write(_xtc_pack(ints, precision))


def save_traj_npz(output_filename, dims, coords, name, mass=None):
    """Write the traj npz exactly as the reference does (load_traj.py:110): typ = atom names."""
    name = np.asarray(name)
    mass = np.zeros(len(name)) if mass is None else np.asarray(mass)
    # same container and keys as np.savez_compressed; deflated on all cores, with a piece index so that
    # NpzFrameStream can inflate ``coords`` on all cores as well
    npz_writer.savez_parallel(output_filename, chunk=TRAJ_PIECE, dims=dims, coords=coords, name=name, mass=mass, typ=name)


def process_gro_mdtraj(topology_filename, trajectory_filename, output_filename):
    """Trajectory + topology -> ``output_filename.npz`` (reference load_traj.py:90-111)."""
    print("processing ", trajectory_filename)
    try:
        import mdtraj as md
    except ImportError as exc:
        if trajectory_filename.endswith(".gro"):
            names, xyz, box = read_gro_frames(trajectory_filename)
        elif trajectory_filename.endswith(".trr"):
            names = load_gro(topology_filename)
            xyz, box, _ = read_trr(trajectory_filename)
            if xyz.shape[1] != len(names):
                raise ValueError("%s has %d atoms, topology %s has %d" % (trajectory_filename, xyz.shape[1], topology_filename, len(names)))
        elif trajectory_filename.endswith(".dcd"):       # the CLI's NAMD route (main_gromacs.py:95-96): foo.psf + foo.dcd
            if topology_filename.endswith(".psf"):
                names, mass = load_psf(topology_filename)
            else:
                names, mass = load_gro(topology_filename), None
            xyz, box = read_dcd(trajectory_filename)
            if xyz.shape[1] != len(names):
                raise ValueError("%s has %d atoms, topology %s has %d" % (trajectory_filename, xyz.shape[1], topology_filename, len(names)))
            print("saving ", output_filename)
            save_traj_npz(output_filename, box, xyz, names, mass)
            print('done saving')
            return
        elif trajectory_filename.endswith(".xtc"):
            names = load_gro(topology_filename)
            xyz, box, _ = read_xtc(trajectory_filename)
            if xyz.shape[1] != len(names):
                raise ValueError("%s has %d atoms, topology %s has %d" % (trajectory_filename, xyz.shape[1], topology_filename, len(names)))
        else:
            raise ImportError("mdtraj is needed to decode %s (.gro, .trr, .xtc and .dcd are parsed natively)" % trajectory_filename) from exc
        print("saving ", output_filename)
        save_traj_npz(output_filename, box, xyz, names)
        print('done saving')
        return
    t = md.load(trajectory_filename, top=topology_filename)
    coords = t.xyz * 10            # nm -> Angstrom, float32
    dims = t.unitcell_lengths * 10
    name = np.array([a.name for a in t.topology.atoms])
    mass = np.array([a.element.mass for a in t.topology.atoms])
    print("saving ", output_filename)
    save_traj_npz(output_filename, dims, coords, name, mass)
    print('done saving')


class XtcFrameStream:
    """Frames of an .xtc file as a frame source for ``dens.compute_sf_stream`` (same interface as NpzFrameStream:
    ``shape``, ``dtype``, ``read_into``, ``skip``), i.e. trajectory -> pinned chunk buffers without the intermediate
    traj npz the reference writes and re-reads (load_traj.py:110, main_gromacs.py:200-202).  The frame headers are
    walked once at open time, which also yields ``dims`` (T, 3) float32 Angstrom -- every box is needed before frame 0
    (dens.py:52) -- and ``times``; ``read_into`` decodes the next frames' xtc3 blocks on all host cores (libmdsf_io)
    straight into the caller's buffer and converts nm -> Angstrom in place (float32 * 10, the reference's
    ``t.xyz * 10``, load_traj.py:98).  The file is memory-mapped, frames are random access."""

    def __init__(self, path, threads=None):
        self.path = path
        self._data = np.memmap(path, dtype=np.uint8, mode="r")
        self._off, natoms, self.dims, self.times = _scan_xtc(self._data, path)
        self.shape, self.dtype = (len(self._off), natoms, 3), np.dtype(np.float32)
        self._threads = int(threads or 0)
        self.position = 0

    def read_into(self, buf):
        if buf.dtype != self.dtype or buf.shape[1:] != self.shape[1:] or not buf.flags.c_contiguous:
            raise ValueError("chunk buffer must be C-contiguous float32 of shape (k, %d, 3)" % self.shape[1])
        want = min(buf.shape[0], self.shape[0] - self.position)
        if want <= 0:
            return 0
        off = np.ascontiguousarray(self._off[self.position:self.position + want])
        _decode_xtc(self._data, off, self.shape[1], buf[:want], self._threads, self.path, self.position)
        np.multiply(buf[:want], np.float32(10), out=buf[:want])
        self.position += want
        return want

    def skip(self, nframes):
        self.position += max(min(int(nframes), self.shape[0] - self.position), 0)

    def close(self):
        self._data = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class NpzFrameStream:
    """Frames of a trajectory npz (reference layout, load_traj.py:110) WITHOUT loading the whole ``coords`` array.

    The reference does ``traj = np.load(...); T = traj['coords']`` (main_gromacs.py:200-202): the complete
    trajectory is inflated into pageable memory before the first frame is used.  Here the ``coords.npy`` zip
    member is inflated incrementally, straight into the caller's (pinned) chunk buffers, so that decoding chunk
    i+1 overlaps the GPU work on chunk i (dens.compute_sf_stream).

    ``shape`` / ``dtype`` describe the full array; ``read_into(buf)`` fills ``buf[:k]`` with the next k frames and
    returns k (0 at the end); ``skip(n)`` drops n frames."""

    def __init__(self, path, key="coords", threads=None):
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
        # Members written by npz_writer.savez_parallel (our load_traj) carry a piece index: the deflate stream is a
        # chain of independently compressed pieces, so they are inflated on all cores straight from the file
        # (os.pread + zlib release the GIL).  Anything else (np.savez_compressed output) is inflated sequentially.
        self._index = npz_writer.read_piece_index(path, key + ".npy")
        self.parallel = self._index is not None
        if self.parallel:
            self._io = _io_lib()
            self._data0 = self._index["raw_size"] - int(np.prod(shape)) * self.dtype.itemsize     # npy header bytes
            self._fd = os.open(path, os.O_RDONLY)
            self._threads = int(threads or 0)
            self._edge = {}                # piece -> inflated bytes of a piece that straddles two reads

    def _read_parallel(self, view):
        """Fill ``view`` (a writable byte view of the caller's chunk buffer) with the member bytes that follow the
        current position: pieces inside the range are inflated straight into it, on all cores, by libmdsf_io."""
        ix = self._index
        lo = self._data0 + self.position * self._frame_bytes
        hi = lo + len(view)
        starts, lens = ix["raw_start"], ix["raw_len"]
        first = max(int(np.searchsorted(starts, lo, side="right")) - 1, 0)
        last = max(int(np.searchsorted(starts, hi - 1, side="right")) - 1, 0)
        base = np.frombuffer(view, dtype=np.uint8).ctypes.data
        jobs, edges = [], []
        for i in range(first, last + 1):
            s0, s1 = starts[i], starts[i] + lens[i]
            if s0 >= lo and s1 <= hi:
                jobs.append((i, base + (s0 - lo)))
            elif i not in self._edge:
                scratch = np.empty(lens[i], dtype=np.uint8)
                self._edge[i] = scratch
                jobs.append((i, scratch.ctypes.data))
            if not (s0 >= lo and s1 <= hi):
                edges.append(i)
        if jobs:
            n = len(jobs)
            off = (ctypes.c_int64 * n)(*[ix["data_offset"] + ix["comp_start"][i] for i, _ in jobs])
            clen = (ctypes.c_int64 * n)(*[ix["comp_len"][i] for i, _ in jobs])
            rlen = (ctypes.c_int64 * n)(*[lens[i] for i, _ in jobs])
            dst = (ctypes.c_void_p * n)(*[p for _, p in jobs])
            rc = self._io.mdsf_io_inflate_pieces(self._fd, n, off, clen, dst, rlen, self._threads)
            if rc:
                raise EOFError("trajectory npz: piece %d is corrupt or disagrees with its index" % jobs[-rc - 1][0])
        out = np.frombuffer(view, dtype=np.uint8)
        for i in edges:
            s0 = starts[i]
            a, b = max(lo, s0), min(hi, s0 + lens[i])
            out[a - lo:b - lo] = self._edge[i][a - s0:b - s0]
        for i in [k for k in self._edge if starts[k] + lens[k] <= hi]:
            del self._edge[i]

    def read_into(self, buf):
        if buf.dtype != self.dtype or buf.shape[1:] != self.shape[1:] or not buf.flags.c_contiguous:
            raise ValueError("chunk buffer must be C-contiguous %s of shape (k, %d, 3)" % (self.dtype, self.shape[1]))
        want = min(buf.shape[0], self.shape[0] - self.position)
        if want <= 0:
            return 0
        view = memoryview(buf[:want].reshape(-1).view(np.uint8))
        if self.parallel:
            self._read_parallel(view)
            self.position += want
            return want
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
        if self.parallel:                  # random access: nothing to inflate
            self.position += max(nframes, 0)
            return
        left = nframes * self._frame_bytes
        while left > 0:
            chunk = self._fh.read(min(left, 16 << 20))
            if not chunk:
                raise EOFError("trajectory npz ends early")
            left -= len(chunk)
        self.position += max(nframes, 0)

    def close(self):
        if self.parallel:
            os.close(self._fd)
            self._edge.clear()
        self._fh.close()
        self._zip.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
