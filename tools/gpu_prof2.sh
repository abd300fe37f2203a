#!/bin/sh
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"splat_zfft" -s 1 -c 1 -f -o gpurun_out/prof_splat python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-100
