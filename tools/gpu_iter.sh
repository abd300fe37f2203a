#!/bin/sh
# usage: gpu_iter.sh [pytest|nopytest] ; runs parity tests then the c2 bench with a few build/env variants
mkdir -p gpurun_out
if [ "$1" != "nopytest" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
run() { echo "== $1"; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('frames/s %.0f  e2e %.0f  stages(ms/step16):'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"; }
run default X=1
run z16 MDSF_RADIX_LOG2_Z=4
run xy8 MDSF_RADIX_LOG2_XY=3
MDSF_NVCC_FLAGS="-DMDSF_PASS_MINBLOCKS=3" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run pass3 X=1
MDSF_NVCC_FLAGS="-DMDSF_PASS_MINBLOCKS=3 -DMDSF_SPLAT_MINBLOCKS=1" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run pass3_splat1_z16 MDSF_RADIX_LOG2_Z=4
run pass3_splat1_z8 X=1
