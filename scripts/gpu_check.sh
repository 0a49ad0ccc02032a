#!/bin/bash
# First-contact GPU check: smoke, GPU parity tests, short benches. Everything under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -40
echo "== bench tiny" ; timeout 300 python bench.py --workload tiny --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
echo "== bench cfg3" ; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 3000 gpurun_out/bench_cfg3.json; tail -5 gpurun_out/bench_cfg3.err
