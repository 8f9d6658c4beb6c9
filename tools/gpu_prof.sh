#!/bin/bash
# full ncu captures (source-level) of the hot kernels of one bench step: ris_ltc3, winner, trace, gbuffer
mkdir -p gpurun_out
for k in ${KERNELS:-ris_ltc3 winner trace gbuffer}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
