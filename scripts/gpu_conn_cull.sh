#!/bin/bash
# k_conn after the lane-parallel tile culling: ms per launch at 8 / 64 chains (DFMA form and FP32 contraction), cfg5 sweep
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
for c in 8 64; do for a in 0 1; do echo "chains $c acc32 $a"; MP_CONN_ACC32=$a timeout 200 python scripts/conn_micro.py $c 10 0 2>&1 | tail -1; done; done
for i in 1 2; do timeout 300 python bench.py --workload cfg5 --no-extra --no-cpu-baseline --ess-sweeps 0 --steps 16 --warmup 4 > $O/cull_cfg5_$i.json 2> $O/cull_cfg5_$i.err; echo "cfg5 rc=$?"; done
python - <<'PY'
import json
for i in (1,2):
    d=json.loads(open(f'gpurun_out/cull_cfg5_{i}.json').read().strip().splitlines()[-1])
    print('cfg5 ms_per_step %.3f'%d['ms_per_step'], d.get('kernel_ms'), d.get('likelihood_evals_per_sec'))
PY
