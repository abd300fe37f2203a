#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or full_grids or c2_frame or c1_real or layouts or thin or c5 or reproduc or million or general" 2>&1 | tail -3
export MDSF_FUSED_YX=0
run zl1 c3 16 MDSF_ZLANE=1
run zl1 c2 64 MDSF_ZLANE=1
run zl1 c1 64 MDSF_ZLANE=1
KRE=splat WL=c3 FR=8 sh tools/gpu_src.sh
