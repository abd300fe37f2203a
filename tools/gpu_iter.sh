#!/bin/sh
mkdir -p gpurun_out
if [ "$1" != "nopytest" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  stages(ms/step):'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['config'].get('splat'))"; }
run scatter_r3 X=1
run scatter_17mb_r2 MDSF_SLAB_MB=17 MDSF_SLAB_RING=2
run scatter_17mb_r3 MDSF_SLAB_MB=17 MDSF_SLAB_RING=3
run scatter_r4 MDSF_SLAB_RING=4
run skip_scatter MDSF_SPLAT_SKIP=32
run skip_zpass MDSF_SPLAT_SKIP=64
run skip_both MDSF_SPLAT_SKIP=96
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1_scatter X=1
