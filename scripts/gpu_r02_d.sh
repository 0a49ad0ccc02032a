#!/bin/bash
# round 2, fourth GPU pass: block-grid tests, then the cfg5 scan with and without blocks in several launch shapes
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_blocks.py -m gpu -q -x > $O/r02d_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r02d_tests.log
tail -30 $O/r02d_tests.log
: > $O/r02d_blocks.log
timeout 600 python scripts/blocks_micro.py cfg5 0 2>&1 | tail -1 | tee -a $O/r02d_blocks.log
for k in 3 4 5; do timeout 600 python scripts/blocks_micro.py cfg5 $k 2>&1 | tail -1 | tee -a $O/r02d_blocks.log; done
for cs in 2 8; do MP_BLK_CS=$cs timeout 600 python scripts/blocks_micro.py cfg5 4 2>&1 | tail -1 | tee -a $O/r02d_blocks.log; done
MP_BLK_TPT=2048 timeout 600 python scripts/blocks_micro.py cfg5 4 2>&1 | tail -1 | tee -a $O/r02d_blocks.log
