#!/bin/sh
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"slab_pipeline" -s 1 -c 1 -f -o gpurun_out/prof_r1d python bench.py --steps 2 --warmup 1 --no-cpu --frames-per-step 16 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
