"""world_size-2 gloo test of the frame-sharding driver (sf_distributed.py) on CPU.  The per-rank
engine is replaced by an oracle-backed stand-in (test infrastructure), so the test covers the host
logic of the N>1 path: shard boundaries, the global mean box, the reduce and the rank-0 npz."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from tests.helpers import ROOT, load_case, sf_errors


class OracleEngine:
    """Same call surface as mdsf_native.Engine for the calls compute_sf_sharded makes."""

    def __init__(self, L, typ, rad, ucell, sres):
        from oracle import dens_oracle as orc
        self.orc = orc
        self.box, self.typ, self.rad, self.ucell = L, np.asarray(typ), rad, ucell
        self.N, self.dr = orc.grid_shape(L, sres)
        self.n = tuple(int(v) for v in self.N)
        self.widths = orc.half_widths(rad, self.dr, set(self.typ.tolist()))
        self.nb = orc.border_cells(self.widths)
        self.sf = np.zeros((self.n[0], self.n[1], self.n[2] // 2 + 1))

    def push_frames(self, coords, scale, wrap_range=None, write_back=False):
        orc = self.orc
        for t in range(coords.shape[0]):
            fr = coords[t]
            for d in range(3):
                fr[:, d] *= scale[t, d].astype(fr.dtype) if fr.dtype == np.float32 and self.box.dtype == np.float32 else scale[t, d]
            fr3 = fr[None]
            orc.wrap_frames(fr3, self.box)
            self.sf += orc.power_spectrum(orc.density_frame(fr3[0], self.typ, self.rad, self.widths, self.N, self.dr, self.nb, self.ucell))

    def read_sf(self):
        return self.sf

    def sync(self):
        pass

    def close(self):
        pass


def _factory(L, typ, rad, ucell, sres, coord_dtype, arith, device=None):
    e = OracleEngine(L, typ, rad, ucell, sres)
    return e, e.N, e.dr, e.nb


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import mdsf_b200
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = load_case("mono_f32")
        mdsf_b200.dens.PRINT_DETAILS = False
        mdsf_b200.distributed.compute_sf_sharded(c["coords"].copy(), c["dims"], c["typ"], os.path.join(outdir, "sf"),
                                                 c["rad"], c["ucell"], c["sres"], engine_factory=_factory)
    finally:
        dist.destroy_process_group()


def test_frame_shard_partitions_every_frame_once():
    import mdsf_b200
    fs = mdsf_b200.distributed.frame_shard
    for n in (0, 1, 3, 8, 17):
        for w in (1, 2, 3, 8):
            spans = [fs(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_two_rank_sharded_run_matches_reference_golden(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(str(tmp_path / "sf.npz"))
    c = load_case("mono_f32")           # 3 frames -> shards of 2 and 1
    assert np.array_equal(z["N"], c["ref_N"]) and np.array_equal(z["L"], c["ref_L"])
    rel, norm = sf_errors(z["sf"], c["ref_sf"])
    assert rel <= 1e-5 and norm <= 1e-12
    assert np.array_equal(z["kgrid"], c["ref_kgrid"])
