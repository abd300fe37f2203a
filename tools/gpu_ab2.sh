#!/bin/sh
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "(golden and tile and native) or full_size_c2 or ragged or reproducible" 2>&1 | tail -3
MDSF_X_ASYNC=1 MDSF_Y_ASYNC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "numpy or full_size_c2 or (golden and tile and native)" 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps ${STEPS:-20} --warmup 4 --no-cpu $EXTRA 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f e2e %.0f ms/step %.3f y %.3f x %.3f splat %.3f prep %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],s['fft_y'],s['fft_x_accum'],s['splat_zfft'],s['prep_bin']))" || tail -5 gpurun_out/err.log; }
run new_default X=1
run noprio MDSF_PREP_PRIO=0
run notwpref MDSF_TW_PREFETCH=0
run old MDSF_TW_PREFETCH=0 MDSF_PREP_PRIO=0
run xasync MDSF_X_ASYNC=1
run yasync MDSF_Y_ASYNC=1
run xyasync MDSF_X_ASYNC=1 MDSF_Y_ASYNC=1
run prep_early MDSF_PREP_EARLY=1
run prep_early_xy MDSF_PREP_EARLY=1 MDSF_X_ASYNC=1 MDSF_Y_ASYNC=1
