#!/bin/bash
# round 2, first GPU pass: tests (incl. the BASELINE-shape parity tests), new bench line, k_conn shapes, chain counts
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nproc > $O/nproc.txt; nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r02a_tests.log
tail -5 $O/r02a_tests.log
B="--steps 10 --warmup 3"
timeout 600 python bench.py $B --weak --no-extra --no-cpu-baseline --ess-sweeps 0 > $O/r02a_weak.json 2> $O/r02a_weak.err; echo "weak rc=$?"
MP_CONN_SHAPE=1 timeout 600 python bench.py $B --weak --no-extra --no-cpu-baseline --ess-sweeps 0 > $O/r02a_weak_shape1.json 2> $O/r02a_weak_shape1.err
MP_CONN_SHAPE=2 timeout 600 python bench.py $B --weak --no-extra --no-cpu-baseline --ess-sweeps 0 > $O/r02a_weak_shape2.json 2> $O/r02a_weak_shape2.err
for c in 16 32; do
  timeout 600 python bench.py $B --chains $c --no-extra --no-cpu-baseline --ess-sweeps 0 > $O/r02a_c$c.json 2> $O/r02a_c$c.err
done
timeout 900 python bench.py $B > $O/r02a_cfg3.json 2> $O/r02a_cfg3.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --ref-seconds 40 > $O/r02a_ref.json 2> $O/r02a_ref.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unparsed', e); continue
    r=d.get('roofline',{})
    print(f, 'value=%.1f'%d.get('value',-1), 'ms=%.3f'%d.get('ms_per_step',-1), 'e2e=%.1f'%d.get('e2e',{}).get('value',-1), 'kernel_ms', d.get('kernel_ms'), 'launch', d.get('kernel_launches'),
          'frac', r.get('frac'), 'exec', (r.get('executed') or {}).get('frac'), 'conn', (r.get('conn') or {}).get('ms_per_launch'), (r.get('conn') or {}).get('frac'))
PY
