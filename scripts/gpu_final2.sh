#!/bin/bash
# Round-end evidence run, part 2: cfg1 lines, driver tests, then one ncu --set full capture of the two dominant kernels
# (after the plain command exited 0).
mkdir -p gpurun_out
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/bench_r01_cfg1.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r01_cfg1.json; echo
timeout 600 python -m pytest tests/test_gpu_driver.py tests/test_gpu_parity.py -q -m gpu -k "driver or exact" --timeout 600 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_y_cull|k_conn" -s 6 -c 2 -f -o gpurun_out/prof_r01_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
