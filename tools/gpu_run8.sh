#!/bin/sh
mkdir -p gpurun_out
export MDSF_FUSED_YX=1
KRE=yx_pass WL=c3 FR=8 sh tools/gpu_src.sh
mv gpurun_out/prof_src_c3.csv gpurun_out/prof_src_yx_c3.csv; mv gpurun_out/prof_src_c3_raw.csv gpurun_out/prof_src_yx_c3_raw.csv
