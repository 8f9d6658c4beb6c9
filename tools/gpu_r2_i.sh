#!/bin/bash
# Round 2, session I: spatial sharing of an SM between the kernels of two overlapped frames -- RIS with fewer warps, winner / trace beside it
mkdir -p gpurun_out
run() {
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'value', round(j['value'],2), {k: round(v,2) for k,v in j['kernels'].items() if k.endswith('_ms')})"
}
run RISLTC_RIS_WARPS=24
run RISLTC_RIS_WARPS=20
run RISLTC_RIS_WARPS=16
run RISLTC_RIS_WARPS=16 RISLTC_WIN_THREADS=256
run RISLTC_RIS_WARPS=12
run RISLTC_RIS_WARPS=12 RISLTC_WIN_THREADS=256
run RISLTC_RIS_WARPS=16 RISLTC_WIN_THREADS=256 RISLTC_TRACE_CTAS=4
run RISLTC_RIS_WARPS=20 RISLTC_WIN_THREADS=256 RISLTC_TRACE_CTAS=5
run RISLTC_WIN_THREADS=256
run RISLTC_TRACE_CTAS=5
