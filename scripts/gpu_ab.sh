#!/bin/bash
mkdir -p gpurun_out
for tpt in 1024 512; do
  echo "== cfg3 TPT=$tpt"; MP_FAST_TPT=$tpt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','kernel_ms')})"
done
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -3
