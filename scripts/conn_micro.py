#!/usr/bin/env python
"""Micro-benchmark of the connectivity kernels on the cfg3 landscape: ms per mp_connectivity call (engine timing, CUDA events)
for k_conn (per-chain parameters) and, when every chain shares (alpha, b), the tensor-core path.
    python scripts/conn_micro.py [chains] [reps] [gemm: 0|1]"""
import os, sys, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import midaspom_b200 as mb
from midaspom_b200 import synth

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
gemm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = synth.make_workload("cfg3")
t = wl["truth"]
eng = mb.Engine(wl["n"], wl["T"], chains, precision=mb.FP32)
eng.set_landscape_coords(wl["px"], wl["py"], wl["area"]); eng.set_source_units(None); eng.set_observations(wl["obs"])
rng = np.random.default_rng(1)
pars = [dict(e=0.3, c=t["c"], alpha=t["alpha"] * (1.0 if gemm else 1.0 + 0.01 * c), b=t["b"]) for c in range(chains)]
eng.set_params(pars)
z = wl["z_true"].astype(np.uint8)
y = np.stack([(z[:-1] & z[1:] & (rng.random((wl["T"] - 1, wl["n"])) < 0.7)).astype(np.uint8) for _ in range(chains)])
eng.set_state(np.stack([z] * chains), y)
eng.connectivity(fetch=False); eng.connectivity(fetch=False)         # the second call culls against the resident S
eng.set_timing(True); eng.get_timing(reset=True); eng.work_counters(reset=True)
for _ in range(reps):
    eng.connectivity(fetch=False)
ms, n = eng.get_timing(reset=True)
w = eng.work_counters()
print(json.dumps(dict(chains=chains, path=eng.conn_path(), contraction="fp64" if os.environ.get("MP_CONN_ACC32", "1") == "0" else "fp32", conn_ms_per_call=ms["conn"] / max(1, n["conn"]), small_ms_per_call=ms["small"] / reps,
                      executed_fraction=(w["conn_exec"] / max(1, w["conn_total"])) if w["conn_total"] else None, gemm_tiles=w["gemm_tiles"])))
