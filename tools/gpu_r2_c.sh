#!/bin/bash
# Round 2, session C: output stage tests, traversal counters, ncu captures of the three big kernels (C2) and of the
# shadow-ray kernel on C3.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c3.json 2> gpurun_out/wl_c3.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c4.json 2> gpurun_out/wl_c4.err
export RISLTC_GBUFFER=raster
for k in ris_ltc3 winner trace4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace4 -s 3 -c 1 -f -o gpurun_out/prof_trace4_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_trace4_c3.log 2>&1
unset RISLTC_GBUFFER
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/bench.err; cat gpurun_out/bench.json
