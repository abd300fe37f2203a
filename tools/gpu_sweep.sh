#!/bin/sh
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f y %.3f x %.3f splat %.3f prep %.3f'%(d['value'],s['fft_y'],s['fft_x_accum'],s['splat_zfft'],s['prep_bin']))"; }
run base5 X=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1 X=1
EXTRA="--workload c3 --frames-per-step 4 --pool 4 --steps 4" run c3 X=1
MDSF_NVCC_FLAGS="-DMDSF_PASS_MINBLOCKS=6" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run mb6 X=1
MDSF_NVCC_FLAGS="-DMDSF_SPLAT_MINBLOCKS=1" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run splat_mb1 X=1
