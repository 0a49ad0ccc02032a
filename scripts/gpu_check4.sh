#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_posterior.py 2>&1 | tail -8
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -8
echo "== bench cfg3" ; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms','kernel_launches','clocks')}); print(d['e2e']); r=d['roofline']; print({k:r[k] for k in ('achieved','peak','frac','share_of_step','ms_per_launch')}, r['conn'])
PY
tail -3 gpurun_out/bench_cfg3.err
echo "== bench cfg2" ; timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2>gpurun_out/bench_cfg2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms','kernel_launches')})
PY
