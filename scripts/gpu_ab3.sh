#!/bin/bash
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -3
echo "== cfg3"; timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','kernel_ms','kernel_launches')}); print(d['e2e']); print({k:(round(v['mean'],5),round(v['rhat'],3),round(v['ess'])) for k,v in d['posterior'].items()})"
echo "== cfg2"; timeout 600 python bench.py --workload cfg2 --steps 32 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','kernel_ms')})"
