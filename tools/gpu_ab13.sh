#!/bin/sh
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps ${STEPS:-20} --warmup 4 --no-cpu $EXTRA 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f e2e %.0f ms/step %.3f y %.3f x %.3f splat %.3f prep %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],s['fft_y'],s['fft_x_accum'],s['splat_zfft'],s['prep_bin']))" || tail -5 gpurun_out/err.log; }
D=$PWD/md-structure-factor_b200
run base X=1
run lut MDSF_LIB=$D/libmdsf_lut.so
run nz MDSF_LIB=$D/libmdsf_nz.so
run base2 X=1
run lut2 MDSF_LIB=$D/libmdsf_lut.so
run nz2 MDSF_LIB=$D/libmdsf_nz.so
