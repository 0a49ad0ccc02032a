#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
echo "== cfg5 x$N"; bash scripts/gpu_cfg5_multi.sh $N cfg5 2>&1 | tail -c 900
echo; echo "== cfg3 x$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_cfg3_x$N.json 2> gpurun_out/bench_cfg3_x$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg3_x$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches','clocks')}); print(d['e2e']); print(d['config']['chains'])
PY
