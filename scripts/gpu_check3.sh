#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_posterior.py 2>&1 | tail -12
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_y_fast|k_conn" -s 6 -c 2 -o gpurun_out/prof_r01a python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
