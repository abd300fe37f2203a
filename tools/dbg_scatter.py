import os, sys
sys.path.insert(0, '.')
import numpy as np
import mdsf_b200
w = __import__("workloads")
wl = w.get("c2")
coords = w.jitter_frames(wl["base"], wl["box"], 2, wl["jitter"], wl["seed0"])
dens = mdsf_b200.dens
def run(splat, nf=1, batch=2):
    eng, n, dr, nb = dens.make_engine(wl["box"], wl["typ"], wl["rad"], wl["ucell"], wl["sres"], np.float32, np.float32, keep_density=True, batch_frames=batch, splat_mode=splat)
    r = coords[:nf].copy()
    eng.push_frames(r, np.ones((nf, 3)), write_back=True); eng.sync()
    d = [eng.debug_density(f) for f in range(nf)]; eng.close(); return d
a = run("owner", 2)
for rep in range(12):
    b = run("scatter", 2)
    for f in range(2):
        d = np.abs(a[f] - b[f])
        bad = np.argwhere(d > 1e-10)
        if len(bad):
            print("rep", rep, "frame", f, "max diff", d.max(), "nbad", len(bad), "first", bad[:6].tolist(), "x planes", sorted(set(bad[:, 0].tolist()))[:20])
print("done")
