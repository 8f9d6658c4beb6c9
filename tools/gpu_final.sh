#!/bin/bash
# Final check of the round at HEAD: parity tests, smoke, both bench arms
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for wl in c3 c4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err
done
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err
python - <<'P'
import json
for f in ['bench', 'bench_reference', 'wl_c3', 'wl_c4']:
    try:
        j = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(j['value'], 3), 'e2e', round(j['e2e']['value'], 3), 'frac', j.get('roofline', {}).get('frac'), {k: round(v, 3) for k, v in j.get('kernels', {}).items() if k.endswith('_ms')})
    except Exception as e:
        print(f, 'failed', e)
P
