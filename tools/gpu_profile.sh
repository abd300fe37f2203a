#!/bin/sh
# round-2 evidence: bench lines (own arm + reference arm), launch lists and ncu --set full captures of every kernel.
# The .ncu-rep files are exported to CSV on the box (raw page) and removed: gpurun_out/ must stay under 64 MiB.
mkdir -p gpurun_out
if [ "${SKIP_BENCH:-0}" != 1 ]; then
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err
tail -c 600 gpurun_out/bench_r02_c3.json; tail -3 gpurun_out/bench_r02_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_c3_reference.json 2> gpurun_out/bench_r02_ref.err
tail -c 300 gpurun_out/bench_r02_c3_reference.json; tail -3 gpurun_out/bench_r02_ref.err
fi
for wl in ${WLS:-c3 c2}; do
  F=$( [ $wl = c2 ] && echo 64 || echo 16 )
  B="python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu --no-extra --frames-per-step $F --pool $F"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$wl.csv $B > gpurun_out/ncu_l.log 2>&1
  timeout 1200 ncu --set full --clock-control none -k regex:'splat|bin_place|bin_count|order_atoms|prep_atoms|fft|yx_pass|tma_pass' -s ${NK:-7} -c ${NK:-7} -o /tmp/prof_r02_$wl -f $B > gpurun_out/ncu_$wl.log 2>&1
  ncu -i /tmp/prof_r02_$wl.ncu-rep --page raw --csv > gpurun_out/prof_r02_${wl}_raw.csv 2>/dev/null
done
ls -la gpurun_out/
