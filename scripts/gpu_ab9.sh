#!/bin/bash
run() { # workload steps warmup
python bench.py --workload $1 --steps $2 --warmup $3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/$2,3) for k,v in d['kernel_ms'].items()})"
}
for cs in 1 2 8; do echo "== cfg3 CS=$cs"; MP_FAST_CS=$cs run cfg3 16 4; done
for v in "4096 8" "4096 16" "8192 16" "2048 8"; do set -- $v; echo "== cfg5t TPT=$1 CS=$2"; MP_FAST_TPT=$1 MP_FAST_CS=$2 run cfg5t 3 3; done
echo "== cfg2"; run cfg2 40 10
echo "== cfg4"; run cfg4 16 4
