#!/bin/sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "cylindrical or plot_grids" 2>&1 | tail -12
timeout 250 python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import mdsf_b200
from oracle import plot2d_oracle as po
rng = np.random.default_rng(5)
n = (86, 86, 83)
L = np.array([86.59, 86.59, 83.43])
D = np.zeros(n + (4,))
for d in range(3):
    v = (np.arange(n[d]) - n[d] / 2) * 2 * np.pi / L[d]
    sh = [1, 1, 1]; sh[d] = n[d]
    D[..., d] = v.reshape(sh)
D[..., 3] = rng.random(n) ** 4
th = 120 * np.pi / 180
ucell = np.array([[1, 0, 0], [np.cos(th), np.sin(th), 0], [0, 0, 1]])
t0 = time.time(); got, _, _ = mdsf_b200.plot2d_gpu.cylindrical_average(D, ucell, fill=False); t1 = time.time()
want, _, _ = po.cylindrical_average(D, ucell, fill=False); t2 = time.time()
ok = np.isfinite(want)
print('c1-size cylindrical average (400 x %d): GPU %.2f s, scipy %.2f s, NaN rings equal %s (%d), max rel diff %.2e' % (n[2], t1 - t0, t2 - t1, np.array_equal(np.isnan(want), np.isnan(got)), (~ok).sum(), np.max(np.abs(got[ok] - want[ok])) / np.abs(want[ok]).max()))
PY
