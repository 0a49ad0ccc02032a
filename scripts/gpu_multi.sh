#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== cfg1 ours"; timeout 300 python bench.py --workload cfg1 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
echo "== cfg1 reference"; timeout 300 python bench.py --workload cfg1 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-260
echo "== cfg3 x$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_cfg3_x$N.json 2> gpurun_out/bench_cfg3_x$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg3_x$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches','clocks')}); print(d['e2e']); print(d['config']['chains'], {k:(round(v['mean'],5),round(v['rhat'],3)) for k,v in d['posterior'].items()})
PY
tail -3 gpurun_out/bench_cfg3_x$N.err
