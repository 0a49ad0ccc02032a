#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r01_cfg3.json 2> gpurun_out/bench_r01_cfg3.err; tail -2 gpurun_out/bench_r01_cfg3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r01_cfg3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ess_per_sec','ess')}); print({k:(round(v['mean'],5),round(v['sd'],5),round(v['ess']),round(v['rhat'],3)) for k,v in d['posterior'].items()})
PY
timeout 600 python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/bench_r01_cfg2.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r01_cfg2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ess_per_sec','ess')})
PY
