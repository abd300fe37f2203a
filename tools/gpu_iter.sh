#!/bin/sh
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  ms/step %.3f stages:'%(d['value'],d['e2e']['value'],d['ms_per_step']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['config'].get('splat'))"; }
run c2 X=1
run c2_early MDSF_PREP_EARLY=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1 X=1
