#!/bin/sh
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "(golden and tile) or full_size_c2 or ragged or reproducible or periodic_fold or general_ucell or two_handles or pinned" 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps ${STEPS:-20} --warmup 4 --no-cpu $EXTRA 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f e2e %.0f ms/step %.3f y %.3f x %.3f splat %.3f prep %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],s['fft_y'],s['fft_x_accum'],s['splat_zfft'],s['prep_bin']))" || tail -5 gpurun_out/err.log; }
run new_default X=1
run sortbin MDSF_DIRECT_BIN=0
run noprio MDSF_PREP_PRIO=0
run noprio_early MDSF_PREP_PRIO=0 MDSF_PREP_EARLY=1
EXTRA="--frames-per-step 64 --pool 64" STEPS=10 run F64 X=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1 X=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1_sortbin MDSF_DIRECT_BIN=0
EXTRA="--workload c3 --frames-per-step 8 --pool 8" STEPS=4 run c3 X=1
