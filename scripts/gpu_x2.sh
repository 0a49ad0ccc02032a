#!/bin/bash
# 2-GPU pass after the k_conn changes: multi-GPU tests (NCCL inside the library), the driver's scaling launch at N=2
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02c_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 $O/r02c_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 > $O/r02_bench_cfg3_x2.json 2> $O/r02_bench_cfg3_x2.err; echo "x2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_cfg3_x2.json').read().strip().splitlines()[-1])
x=(d.get('extra') or {}).get('cfg5_sharded') or {}
print('value=%.1f'%d['value'], 'ms=%.3f'%d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), 'cfg5:', x.get('ms_per_step'), x.get('ranks_hold_identical_draws'), x.get('error'), 'lik', d.get('likelihood_evals_per_sec'))
PY
