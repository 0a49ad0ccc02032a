#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/validate_recovery.py cfg4l 3000 > gpurun_out/recovery_cfg4l.json 2> gpurun_out/recovery_cfg4l.err; tail -2 gpurun_out/recovery_cfg4l.err
timeout 600 python scripts/validate_recovery.py cfg4 2000 > gpurun_out/recovery_cfg4.json 2> gpurun_out/recovery_cfg4.err; tail -2 gpurun_out/recovery_cfg4.err
timeout 600 python scripts/validate_recovery.py cfg2 4000 disperse > gpurun_out/recovery_cfg2.json 2> gpurun_out/recovery_cfg2.err; tail -2 gpurun_out/recovery_cfg2.err
python - <<PY
import json
for w in ('cfg3','cfg4','cfg2'):
    d=json.load(open(f'gpurun_out/recovery_{w}.json'))
    print(w, round(d['seconds'],1), 's', round(d['chain_iters_per_s'],1), {k:round(v,2) for k,v in d['z_score'].items()}, {k:(round(v['rhat'],3), round(v['ess'])) for k,v in d['posterior'].items()})
PY
