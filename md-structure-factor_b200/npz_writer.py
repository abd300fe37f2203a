"""Parallel writer for the reference's ``.npz`` artefacts.

The reference ends ``compute_sf`` with ``np.savez_compressed(out, sf, sfplt, L, N, kgrid, kgridplt)``
(dens.py:346): one zlib stream per array on one core -- 28 s for a 256^3 run whose GPU loop takes
0.2 s, and ~8 GB through zlib at 512^3.  This module writes the SAME container (a zip archive of
``<key>.npy`` members, deflate, readable by ``np.load`` and by the reference's plot2d.py) but
compresses every member as independent chunks on a thread pool (zlib releases the GIL):

* a member's bytes (npy header + C-ordered data) are cut into ``CHUNK`` byte pieces;
* each piece is deflated on its own (raw deflate, level 6 like numpy) and ended with
  ``Z_SYNC_FLUSH`` -- an empty stored block that byte-aligns the stream without setting the
  final-block bit; only the last piece ends with ``Z_FINISH``.  The concatenation is one valid
  deflate stream (pieces never reference each other's history);
* CRC-32 runs over the raw bytes on the calling thread while the workers compress;
* zip64 records are written whenever a size or offset needs them (kgridplt is 4.2 GB at 512^3);
* the compressed size of every piece is recorded in a private zip extra field (``PIECE_INDEX_ID``) of the member's
  local header.  Other zip readers skip unknown extra fields; ``read_piece_index`` / ``load_traj.NpzFrameStream``
  use it to inflate the pieces of a member on all cores as well (trajectory ingest).

Only the standard library and numpy are used; nothing here touches the GPU.
"""
import io
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

CHUNK = 8 << 20          # raw bytes per independently compressed piece
LEVEL = 6                # zlib default, what np.savez_compressed uses
ZIP64_LIMIT = 0xFFFFFFFF
PIECE_INDEX_ID = 0x534D  # "MS": private extra-field id holding the piece index of a member
PIECE_INDEX_MAGIC = b"MDSFPI01"


def _npy_header(arr):
    """The bytes np.save would put in front of ``arr`` (format 1.0/2.0/3.0 as numpy picks)."""
    buf = io.BytesIO()
    d = np.lib.format.header_data_from_array_1_0(arr)
    try:
        np.lib.format.write_array_header_1_0(buf, d)
    except ValueError:                      # header too long for format 1.0
        buf = io.BytesIO()
        np.lib.format.write_array_header_2_0(buf, d)
    return buf.getvalue()


def _deflate_piece(args):
    view, last, level = args
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    return c.compress(view) + c.flush(zlib.Z_FINISH if last else zlib.Z_SYNC_FLUSH)


def _pieces(header, data, chunk):
    """memoryviews over header + data in pieces of ``chunk`` bytes (the header rides with the first piece)."""
    total = len(header) + len(data)
    if total <= chunk:
        yield bytes(header) + bytes(data)
        return
    first = max(chunk - len(header), 0)
    yield bytes(header) + bytes(data[:first])
    for off in range(first, len(data), chunk):
        yield data[off:off + chunk]


def _index_field(first_raw, chunk, csizes):
    """Extra-field record: id, length, magic, raw bytes of the first piece, raw bytes of the others, count, compressed sizes."""
    body = PIECE_INDEX_MAGIC + struct.pack("<QQI", first_raw, chunk, len(csizes)) + struct.pack("<%dQ" % len(csizes), *csizes)
    return struct.pack("<HH", PIECE_INDEX_ID, len(body)) + body


def read_piece_index(path, member):
    """Piece index of ``member`` (e.g. ``"coords.npy"``) of an npz written by ``savez_parallel``.

    Returns None when the member has no index (np.savez_compressed output, stored members), else a dict with
    ``data_offset`` (file offset of the member's first compressed byte), ``raw_size``, and the lists ``raw_start``,
    ``raw_len``, ``comp_start``, ``comp_len`` per piece (raw positions count from the start of the .npy member)."""
    import zipfile
    with zipfile.ZipFile(path) as zf:
        info = zf.getinfo(member)
    if info.compress_type != zipfile.ZIP_DEFLATED:
        return None
    with open(path, "rb") as fh:
        fh.seek(info.header_offset)
        fixed = fh.read(30)
        sig, _, _, _, _, _, _, _, _, nlen, xlen = struct.unpack("<IHHHHHIIIHH", fixed)
        if sig != 0x04034B50:
            return None
        fh.seek(nlen, 1)
        extra = fh.read(xlen)
        data_offset = fh.tell()
    pos = 0
    while pos + 4 <= len(extra):
        fid, flen = struct.unpack_from("<HH", extra, pos)
        body = extra[pos + 4:pos + 4 + flen]
        pos += 4 + flen
        if fid != PIECE_INDEX_ID or not body.startswith(PIECE_INDEX_MAGIC):
            continue
        first_raw, chunk, n = struct.unpack_from("<QQI", body, len(PIECE_INDEX_MAGIC))
        csizes = struct.unpack_from("<%dQ" % n, body, len(PIECE_INDEX_MAGIC) + 20)
        raw_size = info.file_size
        raw_start, raw_len, comp_start, comp_len = [], [], [], []
        r, c = 0, 0
        for i in range(n):
            ln = min(first_raw if i == 0 else chunk, raw_size - r)
            raw_start.append(r); raw_len.append(ln); comp_start.append(c); comp_len.append(csizes[i])
            r += ln; c += csizes[i]
        if r != raw_size or c != info.compress_size:
            return None
        return dict(data_offset=data_offset, raw_size=raw_size, raw_start=raw_start, raw_len=raw_len,
                    comp_start=comp_start, comp_len=comp_len)
    return None


def savez_parallel(file, compressed=True, threads=None, chunk=CHUNK, level=LEVEL, force_zip64=False, **arrays):
    """``np.savez_compressed(file, **arrays)`` with the deflate work spread over ``threads`` cores
    (``compressed=False``: stored members, like ``np.savez``).  ``file`` gets ``.npz`` appended when missing,
    as numpy does."""
    if not isinstance(file, (str, os.PathLike)):
        raise TypeError("savez_parallel writes to a path")
    file = os.fspath(file)
    if not file.endswith(".npz"):
        file += ".npz"
    threads = threads or min(32, os.cpu_count() or 1)
    method = 8 if compressed else 0
    central = []
    with open(file, "wb") as fh, ThreadPoolExecutor(max_workers=threads) as pool:
        for key, val in arrays.items():
            arr = np.asanyarray(val)
            if arr.dtype.hasobject:
                raise TypeError("object arrays are not supported")
            if arr.flags.f_contiguous and not arr.flags.c_contiguous:
                data_arr = arr.T                                   # np.save writes Fortran arrays transposed
            else:
                data_arr = np.ascontiguousarray(arr)
            header = _npy_header(arr)
            data = memoryview(data_arr.reshape(-1).view(np.uint8)) if data_arr.size else memoryview(b"")
            name = (key + ".npy").encode("utf-8")
            raw_size = len(header) + len(data)
            zip64 = force_zip64 or raw_size >= ZIP64_LIMIT or fh.tell() >= ZIP64_LIMIT
            offset = fh.tell()
            mchunk = chunk
            while compressed and 64 + 8 * (raw_size // mchunk + 2) > 60000:     # the index must fit a zip extra field
                mchunk *= 2
            views = list(_pieces(header, data, mchunk))
            indexed = compressed and len(views) > 1
            csizes = [0] * len(views)
            # local header with the sizes still unknown: written again once the member is complete
            extra = (struct.pack("<HHQQ", 1, 16, 0, 0) if zip64 else b"") + (_index_field(len(views[0]), mchunk, csizes) if indexed else b"")
            fh.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 45 if zip64 else 20, 0x800, method, 0, 0x21, 0, 0, 0, len(name), len(extra)))
            fh.write(name)
            fh.write(extra)
            crc, csize = 0, 0
            if compressed:
                jobs = pool.map(_deflate_piece, [(v, i == len(views) - 1, level) for i, v in enumerate(views)])
                for i, (v, blob) in enumerate(zip(views, jobs)):
                    crc = zlib.crc32(v, crc)
                    fh.write(blob)
                    csize += len(blob)
                    csizes[i] = len(blob)
            else:
                for v in views:
                    crc = zlib.crc32(v, crc)
                    fh.write(v)
                    csize += len(v)
            end = fh.tell()
            fh.seek(offset)
            index = _index_field(len(views[0]), mchunk, csizes) if indexed else b""
            if zip64:
                extra = struct.pack("<HHQQ", 1, 16, raw_size, csize) + index
                fh.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 45, 0x800, method, 0, 0x21, crc, ZIP64_LIMIT, ZIP64_LIMIT, len(name), len(extra)))
            else:
                extra = index
                fh.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0x800, method, 0, 0x21, crc, csize, raw_size, len(name), len(extra)))
            fh.write(name)
            fh.write(extra)
            fh.seek(end)
            central.append((name, method, crc, csize, raw_size, offset, zip64))
        # central directory
        cd_start = fh.tell()
        for name, method, crc, csize, raw_size, offset, zip64 in central:
            fields = b""
            if zip64 or offset >= ZIP64_LIMIT:
                fields = struct.pack("<QQQ", raw_size, csize, offset)
                extra = struct.pack("<HH", 1, len(fields)) + fields
                fh.write(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 45, 45, 0x800, method, 0, 0x21, crc, ZIP64_LIMIT, ZIP64_LIMIT,
                                     len(name), len(extra), 0, 0, 0, 0o600 << 16, ZIP64_LIMIT))
            else:
                extra = b""
                fh.write(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 0x800, method, 0, 0x21, crc, csize, raw_size,
                                     len(name), 0, 0, 0, 0, 0o600 << 16, offset))
            fh.write(name)
            fh.write(extra)
        cd_size = fh.tell() - cd_start
        n = len(central)
        if any(c[6] for c in central) or cd_start >= ZIP64_LIMIT or n >= 0xFFFF:
            z64_end = fh.tell()
            fh.write(struct.pack("<IQHHIIQQQQ", 0x06064B50, 44, 45, 45, 0, 0, n, n, cd_size, cd_start))
            fh.write(struct.pack("<IIQI", 0x07064B50, 0, z64_end, 1))
            fh.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, min(n, 0xFFFF), min(n, 0xFFFF), min(cd_size, ZIP64_LIMIT), ZIP64_LIMIT, 0))
        else:
            fh.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, n, n, cd_size, cd_start, 0))
    return file
