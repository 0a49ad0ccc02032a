#!/bin/bash
# the driver's scaling launch at N=8 (cfg3 strong scaling, 64 chains = 8 per GPU; extra.cfg5_sharded through NCCL inside the library)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 > $O/r02_bench_cfg3_x8.json 2> $O/r02_bench_cfg3_x8.err; echo "x8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_cfg3_x8.json').read().strip().splitlines()[-1])
x=(d.get('extra') or {}).get('cfg5_sharded') or {}
print('value=%.1f'%d['value'], 'ms=%.3f'%d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), 'cfg5:', x.get('ms_per_step'), x.get('ranks_hold_identical_draws'), x.get('error'), 'lik', d.get('likelihood_evals_per_sec'))
PY
