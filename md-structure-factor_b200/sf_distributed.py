"""Frame-sharded structure factor across the GPUs of one node.

sf is a plain sum over frames (reference dens.py:318), so the trajectory shards by frames with no
data-path collective: every rank accumulates a partial S(q) on its own GPU and one
``torch.distributed.reduce`` (NCCL over NVLink, fp64 sum) combines them at the end.  The only
global prerequisites are O(T): the mean box (reference dens.py:52) and from it N, dr, Nborder.
torch is plumbing here (process group + the tensor the reduce runs on).
"""
import numpy as np

import dens


def frame_shard(nframes, rank, world):
    """Contiguous block of frames for ``rank``; blocks differ by at most one frame."""
    base, extra = divmod(nframes, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def default_device():
    """GPU of this rank when the caller names none: LOCAL_RANK (torchrun), else torch's current device."""
    import os
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except ImportError:
        pass
    return dens.DEVICE


def reduce_partial_sf(engine, group=None, dst=0):
    """Sum the per-rank partial sf onto ``dst``; returns the numpy array there, None elsewhere."""
    import torch
    import torch.distributed as dist
    shape = (engine.n[0], engine.n[1], engine.n[2] // 2 + 1)
    if dist.get_backend(group) == "nccl":
        # the reduce buffer lives on the ENGINE's GPU: export_sf_kernel writes it, NCCL reads it
        dev = torch.device("cuda", int(getattr(engine, "device", torch.cuda.current_device())))
        torch.cuda.set_device(dev)
        buf = torch.empty(shape, dtype=torch.float64, device=dev)
        engine.sync()
        engine.export_sf_device(buf.data_ptr())
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)
        return buf.cpu().numpy() if dist.get_rank(group) == dst else None
    buf = torch.from_numpy(np.ascontiguousarray(engine.read_sf()))
    dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return buf.numpy() if dist.get_rank(group) == dst else None


def compute_sf_sharded(r, L, typ, out_filename, rad, ucell, Sres, group=None, engine_factory=None, device=None):
    """``dens.compute_sf`` with frames sharded over the ranks of ``group``.

    Every rank passes the FULL ``L`` (T,3) (needed for the global mean box) and either the full
    ``r`` or just its own shard ``r[start:stop]`` (detected by the leading dimension).  Rank 0
    writes ``out_filename + '.npz'``.  ``engine_factory`` exists for the CPU/gloo tests."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dims = np.asarray(L)
    if dims.dtype not in (np.float32, np.float64):
        dims = dims.astype(np.float64)
    nframes = dims.shape[0]
    start, stop = frame_shard(nframes, rank, world)
    mine = r if r.shape[0] == stop - start and r.shape[0] != nframes else r[start:stop]
    if mine.dtype not in (np.float32, np.float64) or not mine.flags.c_contiguous:
        mine = np.ascontiguousarray(mine, dtype=np.float64 if mine.dtype not in (np.float32, np.float64) else mine.dtype)
    arith = np.float32 if (mine.dtype == np.float32 and dims.dtype == np.float32) else np.float64
    Lmean = np.average(dims, axis=0)
    scale = (Lmean / dims).astype(np.float64)
    factory = engine_factory or dens.make_engine
    if device is None and engine_factory is None:
        device = default_device()
    eng, n, dr, nborder = factory(Lmean, typ, rad, ucell, Sres, mine.dtype, arith, device=device)
    try:
        lo, hi = dens._wrapped_atoms(nframes, mine.shape[1])
        if stop > start:
            eng.push_frames(mine, scale[start:stop], (lo, hi), write_back=dens.WRITE_BACK_COORDS)
        sf = reduce_partial_sf(eng, group, 0)
    finally:
        eng.close()
    if rank == 0:
        dens.finish_sf(sf, Lmean, n, out_filename)
    return sf
