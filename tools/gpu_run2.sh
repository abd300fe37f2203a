#!/bin/sh
# remaining tests + ncu: launch list and full captures of the splat / prep / bin kernels on c3 and c2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "c5_grid or thin_long or nccl or streamed or xtc_frame or cli_traj or layouts or monoclinic_pre or million or two_handles or device_pageable or ragged" 2>&1 | tail -30 > gpurun_out/pytest_gpu2.log
tail -4 gpurun_out/pytest_gpu2.log
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --frames-per-step 8 --pool 8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'splat|bin_place|prep_atoms' -s 3 -c 3 -o gpurun_out/prof_c3_splat -f $B > gpurun_out/ncu_c3.log 2>&1
B2="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --frames-per-step 64 --pool 64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'splat' -s 1 -c 1 -o gpurun_out/prof_c2_splat -f $B2 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
