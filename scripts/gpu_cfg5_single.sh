#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, {k:round(v/4,3) for k,v in d['kernel_ms'].items()})
PY
tail -2 gpurun_out/bench_cfg5.err
