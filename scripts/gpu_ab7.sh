#!/bin/bash
echo "== cfg3"; timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/16,3) for k,v in d['kernel_ms'].items()})"
echo "== cfg5t"; timeout 600 python bench.py --workload cfg5t --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/3,3) for k,v in d['kernel_ms'].items()})"
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -4
