#!/bin/bash
# one short bench line per BASELINE.json workload (c1, c3, c4 at 1 GPU; c2 is the default bench)
for wl in c1 c3 c4; do
  SECONDS=0; timeout 900 python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/wl_$wl.err | tail -1 > gpurun_out/wl_$wl.json
  echo "$wl wall $SECONDS s"; tail -n 3 gpurun_out/wl_$wl.err
  python -c "
import json,sys
try:
    j=json.loads(open('gpurun_out/wl_$wl.json').read()); print('$wl', 'value', round(j['value'],2), 'frac', round(j['roofline']['frac'],4), 'ms/step', round(j['ms_per_step'],1), {k: round(v,2) for k,v in j['kernels'].items() if k.endswith('_ms')}, 'e2e', round(j['e2e']['value'],2))
except Exception as e: print('$wl failed', e)"
done
