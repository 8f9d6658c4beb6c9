#!/bin/bash
# One GPU session: parity tests, smoke, bench, launch list, full ncu captures of the kernels of one frame.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# under the profiler the timed choice between the two visibility implementations is distorted by the per-kernel overhead
# (the rasteriser is three kernels and two memsets, the BVH walk one kernel): pin what the un-profiled bench chooses on C2
export RISLTC_GBUFFER=raster
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for k in ${KERNELS:-ris_ltc3 winner trace4 raster_tiles resolve_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
unset RISLTC_GBUFFER
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err; cat gpurun_out/bench.json
