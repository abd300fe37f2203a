#!/bin/sh
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-200 gpurun_out/bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['stage_ms_per_step'])"
python tools/stream_bench.py 1024 --skip-ref 2>&1 | grep -v "^GPU engine\|Calculating\|Progress" | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
