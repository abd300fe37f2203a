#!/bin/sh
# first GPU pass of the rewritten engine: parity suite, then quick bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for wl in c2 c3; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --frames-per-step $( [ $wl = c2 ] && echo 64 || echo 8 ) --pool 16 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  tail -c 1500 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
