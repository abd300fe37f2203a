#!/bin/sh
mkdir -p gpurun_out
for mb in 8 32; do
MDSF_SLAB_MB=$mb ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:"slab_pipeline" -s 1 -c 1 --csv --log-file gpurun_out/pipe_$mb.csv python bench.py --steps 2 --warmup 1 --no-cpu --frames-per-step 16 > /dev/null 2>&1
grep slab_pipeline gpurun_out/pipe_$mb.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
