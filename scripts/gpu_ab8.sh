#!/bin/bash
cp midaspom_b200/lib/libmidaspom_cuda.so /tmp/keep.so
run() {
echo "== cfg3 $1"; timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/16,3) for k,v in d['kernel_ms'].items()})"
echo "== cfg5t $1"; timeout 600 python bench.py --workload cfg5t --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/3,3) for k,v in d['kernel_ms'].items()})"
}
for sp in 2 3; do cp midaspom_b200/lib/libspec$sp.so midaspom_b200/lib/libmidaspom_cuda.so; touch midaspom_b200/lib/libmidaspom_cuda.so; run SP=$sp; done
cp /tmp/keep.so midaspom_b200/lib/libmidaspom_cuda.so
