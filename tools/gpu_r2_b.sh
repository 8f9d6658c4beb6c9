#!/bin/bash
# Round 2, session B: restructured winner kernel (three register budgets, correctly rounded variant), bench lines of every
# workload, the C5 sweep, traversal counters, FFMA2 micro-benchmark.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
for t in 320 256; do
  RISLTC_WIN_THREADS=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_win$t.json 2> gpurun_out/bench_win$t.err
done
RISLTC_WINNER=cr timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_winner_cr.json 2> gpurun_out/bench_winner_cr.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c3.json 2> gpurun_out/wl_c3.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --emulate-stripes 8 > gpurun_out/wl_c3_s8.json 2> gpurun_out/wl_c3_s8.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 2 --emulate-stripes 8 > gpurun_out/wl_c2_s8.json 2> gpurun_out/wl_c2_s8.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/wl_c4.json 2> gpurun_out/wl_c4.err
timeout 900 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/wl_c1.json 2> gpurun_out/wl_c1.err
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/wl_c5.json 2> gpurun_out/wl_c5.err
./tools/microbench/ffma2 > gpurun_out/ffma2.txt 2>&1
tail -n 6 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err; cat gpurun_out/bench.json
