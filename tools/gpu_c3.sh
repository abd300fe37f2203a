#!/bin/sh
# c3 (512^3) pass-geometry A/B
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --workload c3 --frames-per-step 8 --pool 8 --steps 4 --warmup 3 --no-cpu 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f ms/step %.3f y %.3f x %.3f splat %.3f prep %.3f'%(d['value'],d['ms_per_step'],s['fft_y'],s['fft_x_accum'],s['splat_zfft'],s['prep_bin']))" || tail -5 gpurun_out/err.log; }
run base X=1
run thry256 MDSF_THR_Y=256
run thry512 MDSF_THR_Y=512
run thrx512 MDSF_THR_X=512
run thrx128 MDSF_THR_X=128
run wy16 MDSF_WY=16 MDSF_THR_Y=256
