#!/bin/bash
# sweep of the shadow-ray kernel's knobs: per-step trace time (ms) on C2 / C3 / C4
mkdir -p gpurun_out
out=gpurun_out/sweep_trace.txt
: > $out
for wl in c2 c3 c4; do
  for tv in ${TRI_VOTES:-1 4 8 12 16 24}; do
    for rf in ${REFILLS:-6}; do
      RISLTC_TRI_VOTE=$tv RISLTC_REFILL=$rf timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['kernels']; t=j['roofline_trace']
print('$wl tri_vote=$tv refill=$rf trace_ms=%.3f nodes/ray=%.2f tris/ray=%.2f value=%.2f' % (k['trace_ms'], t['node_visits_per_ray'], t['triangle_tests_per_ray'], j['value']))" >> $out
    done
  done
done
cat $out
