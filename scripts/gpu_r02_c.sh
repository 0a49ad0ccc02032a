#!/bin/bash
# round 2, third GPU pass: whole GPU suite, default bench line (cfg3, 64 chains), reference arm
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r02c_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r02c_tests.log
tail -15 $O/r02c_tests.log
timeout 900 python bench.py > $O/r02c_cfg3.json 2> $O/r02c_cfg3.err; echo "cfg3 rc=$?"
tail -c 600 $O/r02c_cfg3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_cfg3.json').read().strip().splitlines()[-1])
r=d.get('roofline',{})
print('value=%.1f'%d.get('value',-1), 'ms=%.3f'%d.get('ms_per_step',-1), 'e2e=%.1f'%d.get('e2e',{}).get('value',-1), 'frac', r.get('frac'), 'conn', (r.get('conn') or {}).get('ms_per_launch'))
print(json.dumps(d.get('extra'))[:3000])
PY
