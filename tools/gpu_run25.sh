#!/bin/sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "nccl" 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 6 --warmup 3 --no-extra > gpurun_out/bench_r02_c3_2gpu.json 2> gpurun_out/bench_r02_c3_2gpu.err
tail -c 1500 gpurun_out/bench_r02_c3_2gpu.json; tail -3 gpurun_out/bench_r02_c3_2gpu.err
