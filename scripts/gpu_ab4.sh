#!/bin/bash
for cfgs in "2 1024" "1 1024" "2 512" "1 512" "2 256" "1 256" "1 128"; do set -- $cfgs
  echo "== cfg2 CS=$1 TPT=$2"; MP_FAST_CS=$1 MP_FAST_TPT=$2 timeout 300 python bench.py --workload cfg2 --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, d['kernel_ms']['sweep_y']/40)
except Exception as e: print('failed', e)"
done
