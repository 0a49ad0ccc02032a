#!/bin/bash
# Round-end evidence run on one GPU, part 1: tests, smoke, bench (both arms, all single-GPU workloads), ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt; nproc > gpurun_out/nproc.txt
echo "== tests"; timeout 2400 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench ours"; timeout 900 python bench.py > gpurun_out/bench_r01_cfg3.json 2> gpurun_out/bench_r01_cfg3.err; tail -c 600 gpurun_out/bench_r01_cfg3.json; echo; tail -2 gpurun_out/bench_r01_cfg3.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; cut -c1-200 gpurun_out/bench_r01_ref.json
echo "== bench cfg2/cfg4/cfg5t/cfg1"; for w in cfg2 cfg4 cfg4l cfg5t cfg5; do timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_r01_$w.json 2>/dev/null; cut -c1-220 gpurun_out/bench_r01_$w.json; echo; done
timeout 300 python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/bench_r01_cfg1.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r01_cfg1.json; echo
timeout 300 python bench.py --workload cfg1 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_cfg1_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r01_cfg1_ref.json; echo
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-120
