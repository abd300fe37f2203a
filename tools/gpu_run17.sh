#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
timeout 200 python -m pytest tests -m gpu -x -q -k "numpy or layouts or full_grids" 2>&1 | tail -3
run pf c3 16 X=1
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-extra --frames-per-step 16 --pool 16"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'tma_pass' -c 4 --csv --log-file gpurun_out/launches17_c3.csv $B > gpurun_out/ncu_l.log 2>&1
python tools/launch_table.py gpurun_out/launches17_c3.csv
