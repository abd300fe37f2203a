#!/bin/sh
# phase-skip sweep of the splat kernel (profiling aid): MDSF_SPLAT_SKIP bits 1 phase B, 2 z FFT, 4 staging, 8 store, 16 list loop
WL=${WL:-c2}; FR=${FR:-32}
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --workload $WL --frames-per-step $FR --pool $FR | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  splat %.3f'%(d['value'], d['stage_ms_per_step']['splat_zfft']))"; }
for s in ${SKIPS:-0 1 2 16 18}; do run "skip=$s" MDSF_SPLAT_SKIP=$s; done
