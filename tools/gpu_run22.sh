#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
timeout 400 python -m pytest tests -m gpu -x -q -k "golden or full_grids or c2_frame or c1_real or layouts or thin or c5 or reproduc or numpy or general or noise" 2>&1 | tail -3
run ilv c3 16 X=1
run ilv0 c3 16 MDSF_ZILV=0
run ilv c2 64 X=1
run ilv c1 64 X=1
run ilv c4 8 X=1
