#!/bin/bash
# 2-GPU pass: multi-GPU tests through the C ABI (NCCL inside the library, driver -G), run BEFORE anything imports torch in the test process
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02k_multi.log 2>&1; echo "multi tests rc=$?"; tail -4 $O/r02k_multi.log
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_multi.py -m gpu -q > $O/r02k_multi2.log 2>&1; echo "blocks+multi rc=$?"; tail -3 $O/r02k_multi2.log
make -C driver > /dev/null 2>&1; driver/midaspom -m 400 -d 100 -s 11 -n 400 -c 8 -i tests/golden/occupancies_example.txt -o /tmp/p.txt -G 0,1 | tail -4
