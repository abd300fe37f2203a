#!/bin/sh
# full GPU suite + source-level profile of the splat on c3 (unfused passes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_r5.log; cat gpurun_out/pytest_gpu_r5.log
export MDSF_FUSED_YX=0
KRE=splat WL=c3 FR=8 sh tools/gpu_src.sh
. tools/gpu_ab.sh
run unf5 c3 16 MDSF_FUSED_YX=0
run fus5 c3 16 MDSF_FUSED_YX=1
