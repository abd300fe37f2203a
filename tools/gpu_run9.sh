#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
export MDSF_FUSED_YX=0
run lw0 c3 16 X=1
run lw8 c3 16 MDSF_LAYOUT_W=8
run lw4 c3 16 MDSF_LAYOUT_W=4
run thr c2 64 X=1
run lw8 c2 64 MDSF_LAYOUT_W=8
run lw0 c4 8 X=1
run lw8 c4 8 MDSF_LAYOUT_W=8
