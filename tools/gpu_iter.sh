#!/bin/sh
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  stages(ms/step):'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['config'].get('splat'))"; }
EXTRA="--workload c3 --frames-per-step 4 --pool 4 --tile 4x4" run c3_4x4 X=1
EXTRA="--workload c3 --frames-per-step 4 --pool 4 --tile 2x8" run c3_2x8 X=1
EXTRA="--workload c3 --frames-per-step 4 --pool 4 --tile 4x8" run c3_4x8 X=1
EXTRA="--workload c3 --frames-per-step 8 --pool 8" run c3_f8 X=1
EXTRA="--workload c3 --frames-per-step 4 --pool 4 --splat scatter" run c3_scatter X=1
EXTRA="--workload c3 --frames-per-step 4 --pool 4 --splat scatter" run c3_scatter17 MDSF_SLAB_MB=17
