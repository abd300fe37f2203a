#!/bin/sh
mkdir -p gpurun_out
export MDSF_FUSED_YX=0
KRE=splat WL=c3 FR=8 sh tools/gpu_src.sh
