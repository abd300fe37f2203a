#!/bin/sh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_r16.log; cat gpurun_out/pytest_gpu_r16.log
WLS="c3" sh tools/gpu_profile.sh
