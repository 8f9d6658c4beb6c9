#!/bin/bash
# Round 2, session H: fp32-libm generic kernel for FAST precision (parity of every variant, estimator timings), frame overlap on one GPU, textured C2
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "variants or quad_over_plane or textured" 2>&1 | tail -4
grep "fast" gpurun_out/parity_log.txt | cut -c1-200
for est in uniform_uniform uniform_cp uniform_area cp_cp; do
  timeout 600 python bench.py --estimator $est --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/est_$est.json 2> gpurun_out/est_$est.err
done
for wl in c2 c3 c4; do
  for ov in 0 1; do
    RISLTC_OVERLAP=$ov timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ov_${wl}_$ov.json 2> gpurun_out/ov_${wl}_$ov.err
  done
done
timeout 600 python bench.py --textured --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_textured.json 2> gpurun_out/bench_textured.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/est_*.json') + glob.glob('gpurun_out/ov_*.json') + ['gpurun_out/bench_textured.json']):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['value'], 3), 'e2e', round(j['e2e']['value'], 3), {k: round(v, 3) for k, v in j.get('kernels', {}).items() if k.endswith('_ms')})
    except Exception as e:
        print(f, 'failed', e)
P
