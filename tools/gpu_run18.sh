#!/bin/sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "plot_grids or dropin or streamed or cli_ or random_noise" 2>&1 | tail -3
timeout 200 python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import mdsf_b200
dens = mdsf_b200.dens
workloads = __import__('workloads')
wl = workloads.get('c2')
coords = workloads.jitter_frames(wl['base'], wl['box'], 2, wl['jitter'], wl['seed0'])
dims = np.repeat(wl['box'][None, :], 2, axis=0)
L = np.average(dims, axis=0)
eng, n, dr, nb = dens.make_engine(L, wl['typ'], wl['rad'], wl['ucell'], wl['sres'], np.float32, np.float32)
eng.push_frames(coords, np.ones((2, 3)), (0, coords.shape[1]), write_back=False)
sf = eng.read_sf()
t0 = time.time(); kaxes, paxes = dens._k_axes(sf.shape, L); g = eng.export_plot_grids(kaxes, paxes); t1 = time.time()
sfplt = dens.get_dplot(sf); kg, kp = dens._k_lattices(sf.shape, L); kp[..., 3] = sfplt; t2 = time.time()
print('256^3 plot grids: GPU %.2f s, numpy %.2f s, equal %s' % (t1 - t0, t2 - t1, np.array_equal(g['kgridplt'], kp) and np.array_equal(g['kgrid'], kg) and np.array_equal(g['sfplt'], sfplt)))
eng.close()
PY
