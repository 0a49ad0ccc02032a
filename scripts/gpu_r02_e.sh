#!/bin/bash
# round 2, fifth pass (2 GPUs): block-grid tests, multi-GPU tests through the C ABI, cfg5 with blocks on 1 and 2 GPUs
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_multi.py -m gpu -q > $O/r02e_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r02e_tests.log
tail -30 $O/r02e_tests.log
: > $O/r02e_blocks.log
for k in 0 3 4 5 6; do timeout 600 python scripts/blocks_micro.py cfg5 $k 2>&1 | tail -1 | tee -a $O/r02e_blocks.log; done
MP_BLK_CS=8 timeout 600 python scripts/blocks_micro.py cfg5 5 2>&1 | tail -1 | tee -a $O/r02e_blocks.log
timeout 900 python bench.py --workload cfg5 --steps 4 --warmup 2 > $O/r02e_cfg5_x1.json 2> $O/r02e_cfg5_x1.err; echo "cfg5 x1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload cfg5 --steps 4 --warmup 2 > $O/r02e_cfg5_x2.json 2> $O/r02e_cfg5_x2.err; echo "cfg5 x2 rc=$?"
tail -c 1500 $O/r02e_cfg5_x2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload cfg5 --steps 4 --warmup 2 --torch-collectives > $O/r02e_cfg5_x2_torch.json 2> $O/r02e_cfg5_x2_torch.err; echo "cfg5 x2 torch rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_cfg5*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value=%.2f'%d['value'], 'ms=%.2f'%d['ms_per_step'], d.get('ranks_hold_identical_draws'), d['config'].get('scan'), d['config'].get('scan_blocks'), d.get('phase_ms_per_sweep'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
