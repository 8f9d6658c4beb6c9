timeout 300 python -m pytest tests -m gpu -x -q -k "room_default and fast" 2>&1 | tail -1
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('EXP', round(j['value'],2), {k: round(v,2) for k,v in j['kernels'].items() if k.endswith('_ms')})"
