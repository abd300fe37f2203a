#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
export MDSF_TMA_CLUSTER=4
timeout 100 python -m pytest tests -m gpu -x -q -k "numpy and 512" 2>&1 | tail -3
run cl4 c3 16 MDSF_TMA_CLUSTER=4
run cl2 c3 16 MDSF_TMA_CLUSTER=2
run cl1 c3 16 MDSF_TMA_CLUSTER=1
