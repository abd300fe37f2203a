#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
run tma3 c3 16 MDSF_TMA_PASS=3
run tma1 c3 16 MDSF_TMA_PASS=1
run tma1e c3 16 MDSF_TMA_PASS=1 MDSF_PREP_EARLY=1
run tma3e c3 16 MDSF_TMA_PASS=3 MDSF_PREP_EARLY=1
run tma3f32 c3 32 MDSF_TMA_PASS=3
