#!/bin/sh
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu $EXTRA | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f y %.3f x %.3f'%(d['value'],s['fft_y'],s['fft_x_accum']))"; }
run base X=1
run w16 MDSF_WY=16 MDSF_WX=16
run w4 MDSF_WY=4 MDSF_WX=4
run w16_t256 MDSF_WY=16 MDSF_WX=16 MDSF_THR_Y=256 MDSF_THR_X=256
MDSF_NVCC_FLAGS="-DMDSF_PASS_THREADS=256 -DMDSF_PASS_MINBLOCKS=2" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run b256_w16 MDSF_WY=16 MDSF_WX=16 MDSF_THR_Y=256 MDSF_THR_X=256
run b256_w8 MDSF_THR_Y=256 MDSF_THR_X=256
MDSF_NVCC_FLAGS="-DMDSF_PASS_MINBLOCKS=3" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run mb3 X=1
MDSF_NVCC_FLAGS="-DMDSF_PASS_MINBLOCKS=5" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run mb5 X=1
run mb5_xy8 MDSF_RADIX_LOG2_XY=3
MDSF_NVCC_FLAGS="-DMDSF_PASS_MINBLOCKS=6" sh md-structure-factor_b200/csrc/build.sh > /dev/null 2>&1
run mb6_xy8 MDSF_RADIX_LOG2_XY=3
