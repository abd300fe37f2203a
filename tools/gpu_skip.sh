#!/bin/sh
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  splat %.3f'%(d['value'], d['stage_ms_per_step']['splat_zfft']))"; }
for s in 0 1 2 16 18 ; do run "skip=$s" MDSF_SPLAT_SKIP=$s; done
