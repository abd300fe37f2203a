#!/bin/sh
mkdir -p gpurun_out
python tools/stream_bench.py 128 2>&1 | grep -v "^GPU engine\|Calculating\|Progress" | tail -6
ncu --set full --clock-control none --import-source on -k regex:"prep_atoms|bin_pairs" -s 4 -c 2 -f -o gpurun_out/prof_r01_prep python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_prep.log 2>&1
ls -la gpurun_out/*.ncu-rep
