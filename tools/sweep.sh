#!/bin/bash
# kernel times of one bench step per value of an environment knob: sweep.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$var', '$v', 'value', round(j['value'],2), 'frac', round(j['roofline']['frac'],4), {k: round(v,2) for k,v in j['kernels'].items() if k.endswith('_ms')})"
done
