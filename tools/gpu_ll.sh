#!/bin/sh
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum --clock-control none -s 300 -c 250 --csv --log-file gpurun_out/ll.csv python bench.py --steps 1 --warmup 2 --no-cpu $EXTRA > /dev/null 2>&1
tail -2 gpurun_out/ll.csv | cut -c1-100
