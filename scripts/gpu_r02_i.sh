#!/bin/bash
# round 2, pass i: cluster size of the scan at 64 chains; full default bench line; reference arm
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="--steps 6 --warmup 3 --no-extra --no-cpu-baseline --ess-sweeps 0"
for cs in 1 2 4; do
  MP_FAST_CS=$cs timeout 600 python bench.py $B > $O/r02i_c64_cs$cs.json 2> $O/r02i_c64_cs$cs.err; echo "cs=$cs rc=$?"
done
MP_FAST_TPT=1024 MP_FAST_CS=2 timeout 600 python bench.py $B > $O/r02i_c64_t1024cs2.json 2> $O/r02i_c64_t1024cs2.err
MP_FAST_TPT=1024 MP_FAST_CS=1 timeout 600 python bench.py $B > $O/r02i_c64_t1024cs1.json 2> $O/r02i_c64_t1024cs1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unparsed', e); continue
    r=d.get('roofline',{})
    print(f, 'value=%.1f'%d.get('value',-1), 'ms=%.3f'%d.get('ms_per_step',-1), 'scan_ms %.3f'%r.get('ms_per_launch'), 'conn', (r.get('conn') or {}).get('ms_per_launch'), r.get('scan_geometry') or d['config'].get('scan'))
PY
