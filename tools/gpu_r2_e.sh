#!/bin/bash
# Round 2, session E: pair tracing (trace4p_kernel) -- parity, then C2 / C3 / C4 with the one-ray kernel beside it
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
for wl in c2 c3 c4; do
  for tk in 8 4; do
    RISLTC_TRACE=$tk timeout 900 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/wl_${wl}_t$tk.json 2> gpurun_out/wl_${wl}_t$tk.err
    python - <<P
import json
try:
    j = json.loads(open('gpurun_out/wl_${wl}_t$tk.json').read().strip().splitlines()[-1])
    t = j['roofline_trace']
    print('$wl trace=$tk value', round(j['value'], 2), {k: round(v, 3) for k, v in j['kernels'].items() if k.endswith('_ms')}, 'grays/s', round(t['grays_per_s'], 2), 'nodes/ray', t['node_visits_per_ray'], 'tris/ray', t['triangle_tests_per_ray'])
except Exception as e:
    print('$wl $tk failed', e)
P
  done
done
