#!/bin/sh
mkdir -p gpurun_out
B="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-extra --frames-per-step 64 --pool 64"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches21_c2.csv $B > gpurun_out/ncu_l.log 2>&1
python tools/launch_table.py gpurun_out/launches21_c2.csv
