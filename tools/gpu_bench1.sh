#!/bin/sh
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -5 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --steps 10 --warmup 3 --no-cpu --fft cufft > gpurun_out/bench_c2_cufft.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2_cufft.json
python bench.py --steps 10 --warmup 3 --no-cpu --workload c1 --frames-per-step 64 --pool 64 > gpurun_out/bench_c1.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c1.json
tail -5 gpurun_out/bench_c2.err
