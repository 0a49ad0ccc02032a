#!/bin/bash
cp midaspom_b200/lib/libmidaspom_cuda.so /tmp/keep.so
run() {
echo "== cfg5t $1"; timeout 600 python bench.py --workload cfg5t --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/3,3) for k,v in d['kernel_ms'].items()})"
}
run SP=4
for sp in 6 8; do cp midaspom_b200/lib/libspec$sp.so midaspom_b200/lib/libmidaspom_cuda.so; touch midaspom_b200/lib/libmidaspom_cuda.so; run SP=$sp; done
cp /tmp/keep.so midaspom_b200/lib/libmidaspom_cuda.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 -k "scan_order or variants_agree or large_landscape" 2>&1 | tail -3
