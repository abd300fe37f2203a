#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or full_grids or c2_frame or c1_real or layouts or thin or c5 or reproduc or million or general" 2>&1 | tail -3
export MDSF_FUSED_YX=0
run walk c3 16 X=1
run walk c2 64 X=1
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-extra --frames-per-step 16 --pool 16"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches7_c3.csv $B > gpurun_out/ncu_l.log 2>&1
python tools/launch_table.py gpurun_out/launches7_c3.csv
