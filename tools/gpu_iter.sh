#!/bin/sh
run() { echo "== $1"; shift; env "$@" timeout 400 python bench.py --steps ${ST:-20} --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  ms/step %.3f stages:'%(d['value'],d['e2e']['value'],d['ms_per_step']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['config'].get('splat'), d['config'].get('fft'))"; }
run c2_base X=1
run c2_interleave MDSF_INTERLEAVE=1
MDSF_INTERLEAVE=1 timeout 900 python -m pytest tests -x -q -m gpu -k "golden and native and tile" 2>&1 | tail -3
ST=4 EXTRA="--workload c3 --frames-per-step 8 --pool 8" run c3_base X=1
