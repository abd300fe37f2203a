#!/bin/sh
run() { echo "== $1"; shift; env "$@" python bench.py --steps 6 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  stages(ms/step16):'%(d['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"; }
for z in 3 4; do
for s in 0 1 2 4 8 16 31 ; do run "zlog2=$z skip=$s" MDSF_RADIX_LOG2_Z=$z MDSF_SPLAT_SKIP=$s; done; done
