#!/bin/sh
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 1 --no-cpu --frames-per-step 4 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"splat_zfft|fft_y|fft_x" -s 3 -c 3 -f -o gpurun_out/prof_r1a python bench.py --steps 2 --warmup 1 --no-cpu --frames-per-step 4 >> gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
ls -la gpurun_out
