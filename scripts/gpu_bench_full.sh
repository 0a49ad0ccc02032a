#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01_cfg3.json 2> gpurun_out/bench_r01_cfg3.err
tail -c 1500 gpurun_out/bench_r01_cfg3.json; tail -3 gpurun_out/bench_r01_cfg3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err
tail -c 1200 gpurun_out/bench_r01_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
