#!/bin/sh
# round-style check: smoke, full default bench (with cpu baseline), reference arm, launch list
mkdir -p gpurun_out
export MDSF_RADIX_LOG2_Z=${MDSF_RADIX_LOG2_Z:-4}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err; cat gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
tail -3 gpurun_out/launches.csv | cut -c1-200
