#!/bin/sh
# overlap / SM-partition experiment: correctness first, then a c2 sweep
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/check_split.py 2>&1 | tail -24
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps ${STEPS:-20} --warmup 4 --no-cpu $EXTRA 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f e2e %.0f ms/step %.3f y %.3f x %.3f splat %.3f prep %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],s['fft_y'],s['fft_x_accum'],s['splat_zfft'],s['prep_bin']))" || tail -5 gpurun_out/err.log; }
run base X=1
run xasync MDSF_X_ASYNC=1
run overlap_plain MDSF_SM_SPLIT=-1
run split96 MDSF_SM_SPLIT=96
run split104 MDSF_SM_SPLIT=104
run split112 MDSF_SM_SPLIT=112
run split120 MDSF_SM_SPLIT=120
run split112_xasync MDSF_SM_SPLIT=112 MDSF_X_ASYNC=1
run split104_xasync MDSF_SM_SPLIT=104 MDSF_X_ASYNC=1
EXTRA="--workload c3 --frames-per-step 8 --pool 8" STEPS=4 run c3_base X=1
EXTRA="--workload c3 --frames-per-step 8 --pool 8" STEPS=4 run c3_split120 MDSF_SM_SPLIT=120
