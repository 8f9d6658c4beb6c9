#!/bin/bash
# Round 2, session D: textured materials, tri_vote tuner; bench lines of C2 / C3 / C4 and the 8-stripe emulation
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c3.json 2> gpurun_out/wl_c3.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c4.json 2> gpurun_out/wl_c4.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --emulate-stripes 8 > gpurun_out/wl_c3_s8.json 2> gpurun_out/wl_c3_s8.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 2 --emulate-stripes 8 > gpurun_out/wl_c2_s8.json 2> gpurun_out/wl_c2_s8.err
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/bench.err; cat gpurun_out/bench.json
