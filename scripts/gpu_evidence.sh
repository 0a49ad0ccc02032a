#!/bin/bash
# round 2, evidence pass on one B200: GPU suite, the bench lines of every workload, the reference arm, ncu launch list and
# ncu --set full captures of the hot kernels (each only after its plain run exited 0)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nproc > $O/nproc.txt; nvidia-smi -L > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r02_tests.log
tail -4 $O/r02_tests.log
timeout 900 python bench.py > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --weak --no-extra --ess-sweeps 0 > $O/r02_bench_cfg3_weak8.json 2> $O/r02_bench_cfg3_weak8.err; echo "weak rc=$?"
for wl in cfg2 cfg4 cfg4l cfg5 cfg1; do
  timeout 900 python bench.py --workload $wl --no-extra > $O/r02_bench_$wl.json 2> $O/r02_bench_$wl.err; echo "$wl rc=$?"
done
timeout 600 python bench.py --workload cfg5 --no-blocks --no-extra --no-cpu-baseline --steps 6 --warmup 2 > $O/r02_bench_cfg5_noblocks.json 2> $O/r02_bench_cfg5_noblocks.err
timeout 300 python bench.py --workload cfg1 --impl reference > $O/r02_bench_cfg1_ref.json 2> $O/r02_bench_cfg1_ref.err
for g in 0 1; do timeout 300 python scripts/conn_micro.py 8 10 $g 2>&1 | tail -1; done | tee $O/r02_conn_micro.log
NB="--steps 2 --warmup 1 --no-extra --no-cpu-baseline --ess-sweeps 0"
timeout 600 python bench.py $NB > $O/r02_plain.log 2>&1 && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_ncu_launches.csv python bench.py $NB > $O/r02_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_y_cull -s 2 -c 1 -o $O/prof_r02_scan python bench.py $NB > $O/r02_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conn -s 2 -c 1 -o $O/prof_r02_kconn python bench.py $NB > $O/r02_ncu_kconn.log 2>&1; echo "ncu k_conn rc=$?"
}
timeout 300 python scripts/conn_micro.py 8 3 1 > $O/r02_plain_gemm.log 2>&1 && {
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conn_gemm -s 1 -c 1 -o $O/prof_r02_gemm python scripts/conn_micro.py 8 3 1 > $O/r02_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
}
timeout 300 python bench.py --workload cfg5 --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_plain_cfg5.log 2>&1 && {
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_y_cull -s 40 -c 1 -o $O/prof_r02_scan_cfg5 python bench.py --workload cfg5 --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_scan_cfg5.log 2>&1; echo "ncu scan cfg5 rc=$?"
}
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unparsed', e); continue
    r=d.get('roofline') or {}
    print(f, 'value=%.3f'%d.get('value',-1), d.get('unit'), 'ms=%.3f'%d.get('ms_per_step',-1), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
