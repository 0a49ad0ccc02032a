#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; WL=${2:-cfg5}
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $WL --steps 4 --warmup 3 > gpurun_out/bench_${WL}_x$N.json 2> gpurun_out/bench_${WL}_x$N.err
tail -c 1500 gpurun_out/bench_${WL}_x$N.json; tail -4 gpurun_out/bench_${WL}_x$N.err
