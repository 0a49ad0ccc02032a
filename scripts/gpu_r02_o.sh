#!/bin/bash
# pass o: far commit groups touch S_lo only -- scan parity tests, cfg3 at 8 and 64 chains, cfg5
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocks.py tests/test_gpu_baseline_shapes.py -m gpu -q -x > $O/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02o_tests.log
B="--steps 10 --warmup 3 --no-extra --no-cpu-baseline --ess-sweeps 0"
timeout 600 python bench.py $B --weak > $O/r02o_weak.json 2> $O/r02o_weak.err
timeout 600 python bench.py $B > $O/r02o_c64.json 2> $O/r02o_c64.err
timeout 600 python bench.py $B --workload cfg5 > $O/r02o_cfg5.json 2> $O/r02o_cfg5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline') or {}
    print(f, 'value=%.2f'%d['value'], 'ms=%.3f'%d['ms_per_step'], 'scan_ms', r.get('ms_per_launch'), 'conn', (r.get('conn') or {}).get('ms_per_launch'))
PY
