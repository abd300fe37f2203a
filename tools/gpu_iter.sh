#!/bin/sh
mkdir -p gpurun_out
if [ "$1" != "nopytest" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  stages(ms/step):'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"; }
run scatter32 X=1
run scatter16 MDSF_SLAB_MB=16
run scatter8 MDSF_SLAB_MB=8
run scatter48 MDSF_SLAB_MB=48
EXTRA="--splat owner" run owner X=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1_scatter X=1
