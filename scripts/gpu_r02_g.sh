#!/bin/bash
# ncu of the windowed scan kernel at cfg3 (8 chains)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/wpc_profile.py 8 > $O/r02g_plain.log 2>&1 || exit 1
tail -1 $O/r02g_plain.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_y_wpc -s 3 -c 1 -o $O/prof_r02_wpc python scripts/wpc_profile.py 8 > $O/r02g_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02g_ncu.log
