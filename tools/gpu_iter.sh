#!/bin/sh
MDSF_FUSE_ZY=1 timeout 600 python -m pytest tests -x -q -m gpu -k "full_size_c2 or fft_sizes or reproducible" 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps ${ST:-20} --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  ms/step %.3f stages:'%(d['value'],d['e2e']['value'],d['ms_per_step']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['config'].get('splat'), d['config'].get('fft'))"; }
run c2_zy_d4 MDSF_FUSE_ZY=1 MDSF_ZY_DELAY=4
run c2_zy_d8 MDSF_FUSE_ZY=1 MDSF_ZY_DELAY=8
run c2_zy_d16 MDSF_FUSE_ZY=1 MDSF_ZY_DELAY=16
run c2_zy_d16_noB MDSF_FUSE_ZY=1 MDSF_ZY_DELAY=16 MDSF_SPLAT_SKIP=16
