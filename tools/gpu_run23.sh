#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
D=$PWD/md-structure-factor_b200
run w16 c3 16 MDSF_LIB=$D/libmdsf_w16.so
run w16 c2 64 MDSF_LIB=$D/libmdsf_w16.so
run w16 c1 64 MDSF_LIB=$D/libmdsf_w16.so
run w16 c4 8 MDSF_LIB=$D/libmdsf_w16.so
