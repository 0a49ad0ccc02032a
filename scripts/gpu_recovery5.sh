#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/validate_recovery.py cfg5 400 > gpurun_out/recovery_cfg5.json 2> gpurun_out/recovery_cfg5.err; tail -2 gpurun_out/recovery_cfg5.err
python - <<PY
import json
d=json.load(open('gpurun_out/recovery_cfg5.json'))
print(round(d['seconds'],1), 's', round(d['chain_iters_per_s'],2), {k:round(v,2) for k,v in d['z_score'].items()}, {k:(round(v['mean'],5), round(v['sd'],5), round(v['rhat'],3), round(v['ess'])) for k,v in d['posterior'].items()}, d['truth'])
PY
