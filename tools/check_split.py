"""GPU check: the overlapped / SM-partitioned pipeline (MDSF_SM_SPLIT) and the cp.async x pass (MDSF_X_ASYNC)
give bitwise the S(q) of the serial pipeline over many batches (tile splat mode is order-independent and the
x passes accumulate in batch order on one stream)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdsf_b200  # noqa: E402

workloads = __import__("workloads")


def run(name, nframes, batch, env):
    for k in ("MDSF_SM_SPLIT", "MDSF_X_ASYNC", "MDSF_Y_ASYNC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    wl = workloads.get(name.split("@")[0])
    if "@" in name:         # name@N: same atoms on an N^3 grid (N = 64 / 256 take the two-stage x pass)
        wl["sres"] = float(wl["box"][0]) / int(name.split("@")[1]) * (1 + 1e-6)
    coords = workloads.jitter_frames(wl["base"], wl["box"], nframes, wl["jitter"], wl["seed0"])
    dims = np.repeat(wl["box"][None, :], nframes, axis=0)
    L = np.average(dims, axis=0)
    eng, n, dr, nb = mdsf_b200.dens.make_engine(L, wl["typ"], wl["rad"], wl["ucell"], wl["sres"], np.float32, np.float32,
                                                device=0, batch_frames=batch)
    try:
        eng.push_frames(coords, np.ones((nframes, 3)), write_back=False)
        eng.sync()
        sf = eng.read_sf()
        info = eng.pipeline
    finally:
        eng.close()
    return sf, info


if __name__ == "__main__":
    ok = True
    for name, nframes, batch in (("tiny", 23, 4), ("tiny@64", 23, 4), ("c1", 70, 16), ("tiny@256", 11, 4)):
        ref, info0 = run(name, nframes, batch, {})
        assert not info0["overlap"]
        for env in ({"MDSF_X_ASYNC": "1"}, {"MDSF_SM_SPLIT": "-1"}, {"MDSF_SM_SPLIT": "112"}, {"MDSF_SM_SPLIT": "96", "MDSF_X_ASYNC": "1", "MDSF_Y_ASYNC": "1"}):
            sf, info = run(name, nframes, batch, env)
            same = np.array_equal(sf, ref)
            ok &= same
            print(name, env, info, "bitwise equal" if same else "DIFFERS max rel %.3e" % np.max(np.abs(sf - ref) / ref))
    print("check_split", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
