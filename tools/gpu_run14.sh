#!/bin/sh
mkdir -p gpurun_out
. tools/gpu_ab.sh
timeout 900 python -m pytest tests -m gpu -x -q -k "numpy or layouts or full_grids or golden" 2>&1 | tail -5
run tma c3 16 X=1
run tma0 c3 16 MDSF_TMA_PASS=0
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-extra --frames-per-step 16 --pool 16"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches14_c3.csv $B > gpurun_out/ncu_l.log 2>&1
python tools/launch_table.py gpurun_out/launches14_c3.csv
