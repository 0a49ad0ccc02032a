#!/bin/bash
# what the driver runs at round end, in its order, on a fresh box
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/reh_ref.json 2> gpurun_out/reh_ref.err; wc -l < gpurun_out/reh_ref.json; cut -c1-160 gpurun_out/reh_ref.json
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/reh_ours.json 2> gpurun_out/reh_ours.err; wc -l < gpurun_out/reh_ours.json; cut -c1-200 gpurun_out/reh_ours.json; tail -2 gpurun_out/reh_ours.err
