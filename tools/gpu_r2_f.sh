#!/bin/bash
# Round 2, session F: device-side BVH build -- parity test, then C4 / C3 / C2 with both builders (build time, trace time)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "device_built or any_hit or alternatives" 2>&1 | tail -15
grep "acceleration structure" gpurun_out/parity_log.txt
for wl in c4 c3 c2; do
  for b in host gpu; do
    RISLTC_BVH_BUILD=$b timeout 900 python bench.py --workload $wl --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/wl_${wl}_bvh_$b.json 2> gpurun_out/wl_${wl}_bvh_$b.err
    python - <<P
import json
try:
    j = json.loads(open('gpurun_out/wl_${wl}_bvh_$b.json').read().strip().splitlines()[-1])
    t = j['roofline_trace']
    print('$wl builder=$b value', round(j['value'], 2), j['config']['acceleration_structure'], {k: round(v, 3) for k, v in j['kernels'].items() if k.endswith('_ms')}, 'nodes/ray', round(t['node_visits_per_ray'], 2), 'tris/ray', round(t['triangle_tests_per_ray'], 2))
except Exception as e:
    print('$wl $b failed', e); print(open('gpurun_out/wl_${wl}_bvh_$b.err').read()[-1500:])
P
  done
done
