#!/bin/sh
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"splat_zfft|fft_y|fft_x" -s 3 -c 3 -f -o gpurun_out/prof_r01_c2 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-400
