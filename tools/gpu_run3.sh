#!/bin/sh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in c2 c3; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --frames-per-step $( [ $wl = c2 ] && echo 64 || echo 8 ) --pool 16 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', round(d['value'],1), 'frames/s', 'frac', round(d['roofline']['frac'],3), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})
PY
  tail -2 gpurun_out/bench_$wl.err
done
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --frames-per-step 8 --pool 8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"splat|bin_place|prep_atoms" -s 3 -c 3 -o gpurun_out/prof_c3_splat -f $B > gpurun_out/ncu_c3.log 2>&1
