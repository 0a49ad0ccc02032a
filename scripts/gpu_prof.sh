#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_y_fast" -s 3 -c 1 -o gpurun_out/prof_r01c python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
