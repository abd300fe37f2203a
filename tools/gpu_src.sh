#!/bin/sh
# source-level ncu capture of the splat kernel on c3 (exported to CSV on the box; the .ncu-rep stays there)
mkdir -p gpurun_out
B="python bench.py --workload ${WL:-c3} --steps 1 --warmup 1 --no-cpu --no-extra --frames-per-step ${FR:-8} --pool ${FR:-8}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KRE:-splat}" -s 1 -c 1 -o /tmp/prof_src -f $B > gpurun_out/ncu_src.log 2>&1
ncu -i /tmp/prof_src.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_src_${WL:-c3}.csv 2>/dev/null
ncu -i /tmp/prof_src.ncu-rep --page raw --csv > gpurun_out/prof_src_${WL:-c3}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_src*
