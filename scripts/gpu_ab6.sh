#!/bin/bash
cp midaspom_b200/lib/libmidaspom_cuda.so /tmp/keep.so
for cull in 0 1; do
  echo "== cfg3 CULL=$cull"; MP_FAST_CULL=$cull timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/16,3) for k,v in d['kernel_ms'].items()})
except Exception as e: print('failed', e)"
done
for cull in 0 1; do
  echo "== cfg5t CULL=$cull"; MP_FAST_CULL=$cull timeout 600 python bench.py --workload cfg5t --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/3,3) for k,v in d['kernel_ms'].items()})
except Exception as e: print('failed', e)"
done
echo "== parity subset"; MP_FAST_CULL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 2>&1 | tail -3
bash scripts/gpu_dbg_cull.sh 2>&1 | tail -4
cp /tmp/keep.so midaspom_b200/lib/libmidaspom_cuda.so
