#!/bin/sh
# A/B of engine builds (MDSF_LIB) / knobs on the bench workloads: prints frames/s and stage times
mkdir -p gpurun_out
run() {  # name env...
  name=$1; shift
  for wl in ${WLS:-c2 c3}; do
    env "$@" timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu --frames-per-step $( [ $wl = c2 ] && echo 64 || echo 8 ) --pool 16 > gpurun_out/ab_${name}_$wl.json 2> gpurun_out/ab_${name}_$wl.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_${name}_$wl.json').read().strip().splitlines()[-1])
    print('$name $wl', round(d['value'],1), 'frames/s', {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})
except Exception as e:
    print('$name $wl FAILED', e)
PY
  done
}
run base X=1
run lw8 MDSF_LAYOUT_W=8
run lw4 MDSF_LAYOUT_W=4
