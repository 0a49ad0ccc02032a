#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg4l --no-cpu-baseline > gpurun_out/bench_r01_cfg4l.json 2> gpurun_out/bench_r01_cfg4l.err; cut -c1-230 gpurun_out/bench_r01_cfg4l.json; tail -2 gpurun_out/bench_r01_cfg4l.err
timeout 900 python scripts/validate_recovery.py cfg4l 3000 > gpurun_out/recovery_cfg4l.json 2> gpurun_out/recovery_cfg4l.err; tail -2 gpurun_out/recovery_cfg4l.err
python - <<PY
import json
d=json.load(open('gpurun_out/recovery_cfg4l.json'))
print(round(d['seconds'],1), 's', round(d['chain_iters_per_s'],1), {k:round(v,2) for k,v in d['z_score'].items()}, {k:(round(v['mean'],4), round(v['sd'],4), round(v['rhat'],3), round(v['ess'])) for k,v in d['posterior'].items()})
PY
