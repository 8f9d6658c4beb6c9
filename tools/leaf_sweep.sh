#!/bin/bash
# kernel times of one bench step per BVH leaf size
for leaf in 4 6 8 12; do
  RISLTC_BVH_LEAF=$leaf timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('leaf', $leaf, 'value', round(j['value'],2), 'frac', round(j['roofline']['frac'],4), {k: round(v,2) for k,v in j['kernels'].items() if k.endswith('_ms')})"
done
