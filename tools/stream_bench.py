"""GPU box: frames/s of the CLI's streamed path (traj npz -> NpzFrameStream -> pinned chunks -> engine), c2 frames.
Compares an indexed traj npz (load_traj.save_traj_npz, inflated on all cores by libmdsf_io) with a plain
np.savez_compressed one (sequential inflate).  The npz write of S(q) is excluded (finish_sf is stubbed)."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdsf_b200  # noqa: E402

workloads = __import__("workloads")
dens, lt = mdsf_b200.dens, mdsf_b200.load_traj

if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    wl = workloads.get("c2")
    pool = workloads.jitter_frames(wl["base"], wl["box"], 64, wl["jitter"], wl["seed0"])
    coords = np.concatenate([pool] * (T // 64))
    dims = np.repeat(wl["box"][None, :], len(coords), axis=0)
    d = tempfile.mkdtemp()
    t = time.perf_counter(); lt.save_traj_npz(os.path.join(d, "par_traj"), dims, coords, wl["typ"]); tw_par = time.perf_counter() - t
    skip_ref = "--skip-ref" in sys.argv
    tw_ref = float("nan")
    if not skip_ref:
        t = time.perf_counter(); np.savez_compressed(os.path.join(d, "ref_traj"), dims=dims, coords=coords, typ=wl["typ"]); tw_ref = time.perf_counter() - t
    print("write %d frames: indexed/parallel %.1f s, np.savez_compressed %.1f s (%d cores)" % (len(coords), tw_par, tw_ref, os.cpu_count()))
    dens.finish_sf = lambda sf, L, n, out: None
    dens.PRINT_DETAILS = False
    for name in (("par_traj", "par_traj") if skip_ref else ("par_traj", "ref_traj", "par_traj")):
        with lt.NpzFrameStream(os.path.join(d, name + ".npz")) as fs:
            t = time.perf_counter()
            dens.compute_sf_stream(fs, dims, wl["typ"], os.path.join(d, "out"), wl["rad"], wl["ucell"], wl["sres"])
            dt = time.perf_counter() - t
            print("%s: %s inflate, %.0f frames/s end to end (%.2f ms/frame, engine create included)"
                  % (name, "parallel" if fs.parallel else "sequential", len(coords) / dt, 1e3 * dt / len(coords)))
