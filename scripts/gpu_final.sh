#!/bin/bash
# Round-end evidence run on one GPU: tests, bench (both arms), ncu launch list, ncu full capture.
mkdir -p gpurun_out
echo "== tests"; timeout 2400 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench ours"; timeout 900 python bench.py > gpurun_out/bench_r01_cfg3.json 2> gpurun_out/bench_r01_cfg3.err; tail -c 600 gpurun_out/bench_r01_cfg3.json; echo; tail -2 gpurun_out/bench_r01_cfg3.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; cut -c1-200 gpurun_out/bench_r01_ref.json
echo "== bench cfg2/cfg4/cfg1"; for w in cfg2 cfg4; do timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_r01_$w.json 2>/dev/null; cut -c1-220 gpurun_out/bench_r01_$w.json; echo; done
timeout 300 python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/bench_r01_cfg1.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r01_cfg1.json; echo
timeout 300 python bench.py --workload cfg1 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_cfg1_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r01_cfg1_ref.json; echo
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-120
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_y_fast|k_conn" -s 6 -c 2 -o gpurun_out/prof_r01_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
