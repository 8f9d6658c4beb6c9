#!/bin/bash
# parity tests + one bench line (kernel times) -- the quick loop while tuning a kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_quick.json
python -c "
import json; j=json.load(open('gpurun_out/bench_quick.json')); print('value', round(j['value'],2), 'frac', round(j['roofline']['frac'],4), 'e2e', round(j['e2e']['value'],2), {k: round(v,2) for k,v in j['kernels'].items() if k.endswith('_ms')})"
