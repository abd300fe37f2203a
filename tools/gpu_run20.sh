#!/bin/sh
mkdir -p gpurun_out
run2() {  # name workload frames extra-args
  name=$1; wl=$2; fr=$3; shift 3
  timeout 150 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu --no-extra --frames-per-step $fr "$@" > gpurun_out/ab_${name}_$wl.json 2> gpurun_out/ab_${name}_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_${name}_$wl.json').read().strip().splitlines()[-1])
    print('$name $wl F=$fr', round(d['value'],1), 'frames/s', {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})
except Exception as e:
    print('$name $wl FAILED', e); print(open('gpurun_out/ab_${name}_$wl.err').read()[-600:])
PY
}
run2 nat c3 16
run2 cuf c3 16 --fft cufft
run2 nat c3s 16
run2 cuf c3s 16 --fft cufft
