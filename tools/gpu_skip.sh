#!/bin/sh
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  stages(ms/step16):'%(d['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"; }
for s in 0 32 64 96 ; do run "skip=$s" MDSF_SPLAT_SKIP=$s; done
