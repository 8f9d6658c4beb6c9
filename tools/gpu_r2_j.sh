#!/bin/bash
# Round 2, session J: PLOC device builder -- parity test, then build time and traversal cost of host / PLOC / radix trees on C4, C3, C2
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "device_built" 2>&1 | tail -12
grep "acceleration structure" gpurun_out/parity_log.txt
for wl in c4 c3 c2; do
  for b in host gpu radix; do
    RISLTC_BVH_BUILD=$b timeout 900 python bench.py --workload $wl --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/wl_${wl}_bvh_$b.json 2> gpurun_out/wl_${wl}_bvh_$b.err
    python - <<P
import json
try:
    j = json.loads(open('gpurun_out/wl_${wl}_bvh_$b.json').read().strip().splitlines()[-1])
    t = j['roofline_trace']; a = j['config']['acceleration_structure']
    print('$wl builder=$b value', round(j['value'], 2), 'build', a['build_ms'], a['device_ms'], 'depth', a['depth'], {k: round(v, 2) for k, v in j['kernels'].items() if k.endswith('_ms')}, 'nodes/ray', round(t['node_visits_per_ray'], 2), 'tris/ray', round(t['triangle_tests_per_ray'], 2))
except Exception as e:
    print('$wl $b failed', e); print(open('gpurun_out/wl_${wl}_bvh_$b.err').read()[-1500:])
P
  done
done
