#!/bin/bash
# round 2, sixth pass: the windowed one-CTA scan (mp_sweep_wpc.cuh): whole GPU suite, then cfg3 timings per window
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02f_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r02f_tests.log
tail -40 $O/r02f_tests.log
B="--steps 10 --warmup 3 --no-extra --no-cpu-baseline --ess-sweeps 0"
for w in 8 16 4; do
  MP_WPC_WINDOW=$w timeout 600 python bench.py $B --weak > $O/r02f_weak_w$w.json 2> $O/r02f_weak_w$w.err; echo "weak w=$w rc=$?"
done
MP_WPC=0 timeout 600 python bench.py $B --weak > $O/r02f_weak_off.json 2> $O/r02f_weak_off.err
timeout 600 python bench.py $B > $O/r02f_c64.json 2> $O/r02f_c64.err; echo "c64 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unparsed', e); continue
    r=d.get('roofline',{})
    print(f, 'value=%.1f'%d.get('value',-1), 'ms=%.3f'%d.get('ms_per_step',-1), 'e2e=%.1f'%d.get('e2e',{}).get('value',-1), 'scan_ms', r.get('ms_per_launch'),
          'frac', r.get('frac'), 'exec', (r.get('executed') or {}).get('frac'), 'conn', (r.get('conn') or {}).get('ms_per_launch'), d['config'].get('scan'))
PY
timeout 300 python scripts/blocks_micro.py cfg5 6 2>&1 | tail -1 | tee $O/r02f_blocks.log
timeout 300 python scripts/blocks_micro.py cfg5 7 2>&1 | tail -1 | tee -a $O/r02f_blocks.log
