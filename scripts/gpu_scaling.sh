#!/bin/bash
# 8-GPU box: the driver's scaling launch at N=8 and N=4 (cfg3 strong scaling, 64 chains; extra.cfg5_sharded through NCCL inside the library), multi-GPU tests
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02_x8_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 $O/r02_x8_tests.log
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n > $O/r02_bench_cfg3_x$n.json 2> $O/r02_bench_cfg3_x$n.err; echo "x$n rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --workload cfg5 --steps 8 --warmup 3 --blocks-k 6 > $O/r02_bench_cfg5_x8_k6.json 2> $O/r02_bench_cfg5_x8_k6.err; echo "cfg5 x8 k6 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*_x*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsed',e); continue
    x=(d.get('extra') or {}).get('cfg5_sharded') or {}
    print(f, 'value=%.1f'%d['value'], 'ms=%.3f'%d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), 'cfg5:', x.get('ms_per_step'), x.get('ranks_hold_identical_draws'), x.get('error'), d.get('ranks_hold_identical_draws'))
PY
