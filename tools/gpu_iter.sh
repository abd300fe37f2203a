#!/bin/sh
# usage: gpu_iter.sh [pytest|nopytest] ; runs parity tests then the c2 bench with a few env variants
mkdir -p gpurun_out
if [ "$1" != "nopytest" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
run() { echo "== $1"; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  stages(ms/step16):'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"; }
run default X=1
run z16 MDSF_RADIX_LOG2_Z=4
EXTRA="--tile 2x4" run tile2x4 X=1
EXTRA="--workload c1 --frames-per-step 64 --pool 64" run c1 X=1
