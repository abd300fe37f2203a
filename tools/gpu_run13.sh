#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
export MDSF_FUSED_YX=0
D=$PWD/md-structure-factor_b200
run w12r32 c3 16 X=1
run w12r8 c3 16 MDSF_LIB=$D/libmdsf_w12r8.so
run w8r8 c3 16 MDSF_LIB=$D/libmdsf_w8.so
run w8r8 c2 64 MDSF_LIB=$D/libmdsf_w8.so
run w8r8 c1 64 MDSF_LIB=$D/libmdsf_w8.so
run w8r8zl0 c3 16 MDSF_LIB=$D/libmdsf_w8.so MDSF_ZLANE=0
run w8r8 c4 8 MDSF_LIB=$D/libmdsf_w8.so
MDSF_LIB=$D/libmdsf_w8.so timeout 600 python -m pytest tests -m gpu -x -q -k "golden or full_grids or c1_real or thin" 2>&1 | tail -3
