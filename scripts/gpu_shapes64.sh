#!/bin/bash
# k_conn CTA shapes at 64 chains (cfg3): DFMA form (sampler) and FP32 contraction (evaluation calls)
cd "$(dirname "$0")/.." || exit 1
for s in 1 2 3 4; do echo "dfma shape $s"; MP_CONN_ACC32=0 MP_CONN_SHAPE=$s timeout 200 python scripts/conn_micro.py 64 10 0 2>&1 | tail -1; done
for s in 1 2 3; do echo "fp32 shape $s"; MP_CONN_ACC32=1 MP_CONN_SHAPE=$s timeout 200 python scripts/conn_micro.py 64 10 0 2>&1 | tail -1; done
