#!/bin/sh
if [ "$1" != "nopytest" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  stages(ms/step):'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['config'].get('splat'))"; }
run c2 X=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1 X=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1_nostage MDSF_ZSTAGE=1023
EXTRA="--workload c3 --frames-per-step 4 --pool 4 --steps 4" run c3 X=1
EXTRA="--workload c4 --frames-per-step 4 --pool 4 --steps 4" run c4 X=1
