#!/bin/bash
for cfgs in "4 1024" "2 512" "2 1024" "1 512" "1 1024"; do set -- $cfgs
  echo "== cfg3 CS=$1 TPT=$2"; MP_FAST_CS=$1 MP_FAST_TPT=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','kernel_ms')})
except Exception as e: print('failed', e)"
done
