#!/bin/bash
# one ncu --set full capture of the dominant kernel (after the same command has exited 0 without ncu)
mkdir -p gpurun_out
KREGEX=${KREGEX:-k_sweep_y_cull}
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 3 -c 1 -f -o gpurun_out/prof_r01d python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v/16,3) for k,v in d['kernel_ms'].items()})"
tail -2 gpurun_out/ncu.log
