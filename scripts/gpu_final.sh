#!/bin/bash
# round 2, final pass on one B200 after the k_conn changes (FP32 contraction for the evaluation calls, 32 x 1 shape, lane-parallel
# tile culling): GPU suite, smoke, the bench lines of every single-GPU workload, the ncu launch list of the headline command
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r02c_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r02c_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02c_smoke.log
timeout 900 python bench.py > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --weak --no-extra --ess-sweeps 0 > $O/r02_bench_cfg3_weak8.json 2> $O/r02_bench_cfg3_weak8.err; echo "weak rc=$?"
for wl in cfg2 cfg4 cfg4l cfg5; do
  timeout 900 python bench.py --workload $wl --no-extra > $O/r02_bench_$wl.json 2> $O/r02_bench_$wl.err; echo "$wl rc=$?"
done
NB="--steps 2 --warmup 1 --no-extra --no-cpu-baseline --ess-sweeps 0"
timeout 600 python bench.py $NB > $O/r02_plain.log 2>&1 && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_ncu_launches.csv python bench.py $NB > $O/r02_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
}
for c in 8 64; do for a in 0 1; do MP_CONN_ACC32=$a timeout 200 python scripts/conn_micro.py $c 10 0 2>/dev/null | tail -1; done; done > $O/r02_conn_micro_final.jsonl
timeout 200 python scripts/conn_micro.py 8 10 1 2>/dev/null | tail -1 >> $O/r02_conn_micro_final.jsonl
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unparsed', e); continue
    r=d.get('roofline') or {}
    print(f, 'value=%.3f'%d.get('value',-1), d.get('unit'), 'ms=%.3f'%d.get('ms_per_step',-1), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'lik', d.get('likelihood_evals_per_sec'))
PY
