#!/bin/bash
# quick loop: parity tests + bench lines of C2 (winner CTA sizes), C3, C4
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
summ() { python -c "
import json,sys
j=json.loads(open('$1').read()); k=j['kernels']; t=j['roofline_trace']
print('$1', 'value=%.2f e2e=%.2f frac=%.4f' % (j['value'], j['e2e']['value'], j['roofline']['frac']), {a: round(b,3) for a,b in k.items() if a.endswith('_ms')}, 'Grays/s=%.2f' % t['grays_per_s'])"; }
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; summ gpurun_out/bench.json
for t in ${WIN_THREADS:-}; do RISLTC_WIN_THREADS=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_win$t.json 2> gpurun_out/bench_win$t.err; summ gpurun_out/bench_win$t.json; done
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c3.json 2> gpurun_out/wl_c3.err; summ gpurun_out/wl_c3.json
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c4.json 2> gpurun_out/wl_c4.err; summ gpurun_out/wl_c4.json
