#!/bin/bash
# k_conn FP32 contraction (evaluation entry points): GPU suite, timing against the DFMA form, one ncu --set full capture
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r02b_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r02b_tests.log
timeout 300 python scripts/conn_a32_check.py 10 > $O/r02_conn_a32.json 2> $O/r02_conn_a32.err; echo "a32 rc=$?"; cat $O/r02_conn_a32.json; tail -3 $O/r02_conn_a32.err
timeout 300 python scripts/conn_micro.py 64 3 0 > $O/r02b_plain_conn32.log 2>&1 && {
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conn -s 2 -c 1 -o $O/prof_r02_kconn32 python scripts/conn_micro.py 64 3 0 > $O/r02b_ncu_kconn32.log 2>&1; echo "ncu k_conn32 rc=$?"
}
cat $O/r02b_plain_conn32.log | tail -1
