#!/bin/bash
# k_conn with one target per thread (32 threads x 1 target) where the grid is small: 8 chains at cfg3, one chain at cfg5
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for s in 3 4; do echo "conn_micro shape $s"; MP_CONN_ACC32=0 MP_CONN_SHAPE=$s timeout 200 python scripts/conn_micro.py 8 20 0 2>&1 | tail -1; done
for s in 0 4; do echo "weak8 shape $s"; MP_CONN_SHAPE=$s timeout 300 python bench.py --weak --no-extra --no-cpu-baseline --ess-sweeps 0 --steps 20 --warmup 5 > $O/shape${s}_weak8.json 2> $O/shape${s}_weak8.err; echo "rc=$?"; done
for s in 0 4; do echo "cfg5 shape $s"; MP_CONN_SHAPE=$s timeout 300 python bench.py --workload cfg5 --no-extra --no-cpu-baseline --ess-sweeps 0 --steps 6 --warmup 2 > $O/shape${s}_cfg5.json 2> $O/shape${s}_cfg5.err; echo "rc=$?"; done
python - <<'PY'
import json
for f in ('shape0_weak8','shape4_weak8','shape0_cfg5','shape4_cfg5'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value %.2f ms %.3f e2e %.2f'%(d['value'],d['ms_per_step'],d['e2e']['value']),d.get('kernel_ms'),d.get('kernel_launches',{}).get('conn'))
    except Exception as e: print(f,'ERR',e)
PY
