#!/bin/bash
# 8-GPU pass: threads per block task of the cfg5 scan when each GPU holds only a few (year, block) tasks per colour
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for tpt in 2048 4096; do
  MP_BLK_TPT=$tpt timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800 + tpt / 1024)) bench.py --gpus 8 --workload cfg5 --steps 8 --warmup 3 > $O/r02l_cfg5_x8_t$tpt.json 2> $O/r02l_cfg5_x8_t$tpt.err; echo "tpt $tpt rc=$?"
done
MP_BLK_TPT=2048 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29810 bench.py --gpus 8 --workload cfg5 --steps 8 --warmup 3 --blocks-k 6 > $O/r02l_cfg5_x8_t2048_k6.json 2> $O/r02l_cfg5_x8_t2048_k6.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02l_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'ms=%.3f'%d['ms_per_step'], d['config'].get('scan'), d['config'].get('scan_blocks'), d.get('ranks_hold_identical_draws'))
    except Exception as e:
        print(f,'unparsed',e)
PY
