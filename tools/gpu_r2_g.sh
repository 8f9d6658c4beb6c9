#!/bin/bash
# Round 2, session G: the full record of the round -- parity tests, smoke, bench (both arms), every workload, the five
# estimators of the reference's timing experiment, launch list, full ncu captures of the kernels of one C2 frame.
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for wl in c1 c3 c4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err
done
timeout 1200 python bench.py --workload c5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/wl_c5.json 2> gpurun_out/wl_c5.err
for est in uniform_uniform uniform_cp uniform_area cp_cp; do
  timeout 600 python bench.py --estimator $est --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/est_$est.json 2> gpurun_out/est_$est.err
done
timeout 600 python bench.py --textured --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_textured.json 2> gpurun_out/bench_textured.err
timeout 600 python bench.py --light-vertices 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quads.json 2> gpurun_out/bench_quads.err
RISLTC_OVERLAP=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_overlap0.json 2> gpurun_out/bench_overlap0.err
# under the profiler the timed choice between the two visibility implementations is distorted by the per-kernel overhead: pin what the un-profiled bench chooses on C2
export RISLTC_GBUFFER=raster
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for k in ${KERNELS:-ris_ltc3 winner trace4p raster_tiles resolve_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
unset RISLTC_GBUFFER
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench.json') + glob.glob('gpurun_out/wl_c[1-4].json') + glob.glob('gpurun_out/est_*.json') + glob.glob('gpurun_out/bench_overlap0.json') + glob.glob('gpurun_out/bench_textured.json') + glob.glob('gpurun_out/bench_quads.json') + glob.glob('gpurun_out/bench_reference.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['value'], 3), 'e2e', round(j['e2e']['value'], 3), 'frac', j.get('roofline', {}).get('frac'), {k: round(v, 3) for k, v in j.get('kernels', {}).items() if k.endswith('_ms')})
    except Exception as e:
        print(f, 'failed', e)
try:
    j = json.loads(open('gpurun_out/wl_c5.json').read().strip().splitlines()[-1])
    for s in j['sweep']: print(s)
except Exception as e:
    print('c5 failed', e)
P
