#!/bin/sh
timeout 900 python -m pytest tests -x -q -m gpu -k "fft or full_size_prop" 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps ${ST:-4} --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  ms/step %.3f stages:'%(d['value'],d['ms_per_step']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items() if k in ('fft_y','fft_x_accum','splat_zfft')})"; }
EXTRA="--workload c3 --frames-per-step 8 --pool 8" run c3 X=1
EXTRA="--workload c4 --frames-per-step 4 --pool 4" run c4 X=1
EXTRA="--workload c4 --frames-per-step 4 --pool 4" run c4_x256 MDSF_THR_X=256
EXTRA="--workload c5 --frames-per-step 2 --pool 2" run c5 X=1
ST=20 EXTRA="" run c2 X=1
