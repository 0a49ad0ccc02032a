#!/bin/bash
# round 2, second GPU pass: full test suite (no -x), k_conn micro-benchmark per CTA shape, ncu of k_conn, then the tcgen05 path
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py > $O/r02b_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r02b_tests.log
tail -15 $O/r02b_tests.log
for sh in 1 2; do MP_CONN_SHAPE=$sh timeout 300 python scripts/conn_micro.py 8 10 0; done 2>&1 | tee $O/r02b_conn_micro.log
MP_CONN_SHAPE=2 timeout 300 python scripts/conn_micro.py 64 5 0 2>&1 | tee -a $O/r02b_conn_micro.log
MP_CONN_SHAPE=1 timeout 300 python scripts/conn_micro.py 64 5 0 2>&1 | tee -a $O/r02b_conn_micro.log
MP_CONN_GEMM=0 timeout 300 python scripts/conn_micro.py 8 10 1 2>&1 | tee -a $O/r02b_conn_micro.log
timeout 300 python scripts/conn_micro.py 8 3 0 > $O/plain_conn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conn -s 2 -c 2 -o $O/prof_r02_conn python scripts/conn_micro.py 8 3 0 > $O/ncu_conn.log 2>&1
echo "ncu rc=$?"
# the tensor-core path last (a protocol bug traps the kernel; nothing after it in this call)
timeout 300 python scripts/conn_micro.py 8 10 1 2>&1 | tail -3 | tee $O/r02b_gemm_micro.log
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q > $O/r02b_gemm_tests.log 2>&1; echo "gemm tests rc=$?"
tail -30 $O/r02b_gemm_tests.log
