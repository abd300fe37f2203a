#!/bin/sh
# round-end style validation: full GPU test suite, smoke, ncu captures, default bench, reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:"splat_zfft|fft_y|fft_x" -s 3 -c 3 -f -o gpurun_out/prof_r01_c2 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-600 gpurun_out/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err; cat gpurun_out/bench_reference.json
for t in 2x4 4x2 2x8 8x2; do echo "== tile $t"; python bench.py --steps 10 --warmup 3 --no-cpu --tile $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms_per_step']; print('frames/s %.0f ms/step %.3f y %.3f x %.3f splat %.3f prep %.3f'%(d['value'],d['ms_per_step'],s['fft_y'],s['fft_x_accum'],s['splat_zfft'],s['prep_bin']))"; done
