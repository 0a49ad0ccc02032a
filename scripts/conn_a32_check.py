#!/usr/bin/env python
"""k_conn of the FP32 engines, FP32 contraction (FFMA2 + FP32 partial sums per 32 sources, mp_conn32.cu) against the DFMA form
(MP_CONN_ACC32=0) on the cfg3 landscape: largest relative difference of S, and ms per launch of each at 8 and 64 chains.
    python scripts/conn_a32_check.py [reps]"""
import os, sys, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import midaspom_b200 as mb
from midaspom_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
wl = synth.make_workload("cfg3")
t = wl["truth"]
z = wl["z_true"].astype(np.uint8)


def run(chains, acc32, fetch):
    os.environ["MP_CONN_ACC32"] = str(acc32)                         # read by mp_create
    rng = np.random.default_rng(1)
    eng = mb.Engine(wl["n"], wl["T"], chains, precision=mb.FP32)
    eng.set_landscape_coords(wl["px"], wl["py"], wl["area"]); eng.set_source_units(None); eng.set_observations(wl["obs"])
    eng.set_params([dict(e=0.3, c=t["c"], alpha=t["alpha"] * (1.0 + 0.01 * c), b=t["b"]) for c in range(chains)])
    y = np.stack([(z[:-1] & z[1:] & (rng.random((wl["T"] - 1, wl["n"])) < 0.7)).astype(np.uint8) for _ in range(chains)])
    eng.set_state(np.stack([z] * chains), y)
    S0 = eng.connectivity(fetch=fetch)                               # unculled (no resident S yet)
    S1 = eng.connectivity(fetch=fetch)                               # culled against the resident S
    eng.set_timing(True); eng.get_timing(reset=True)
    for _ in range(reps):
        eng.connectivity(fetch=False)
    ms, n = eng.get_timing(reset=True)
    eng.close()
    return S0, S1, ms["conn"] / max(1, n["conn"])


out = {}
ref0, ref1, ms_d8 = run(8, 0, True)
a0, a1, ms_a8 = run(8, 1, True)
rel = lambda g, w: float(np.max(np.abs(g - w) / np.maximum(np.abs(w), 1e-300)))
out["rel_diff_unculled"] = rel(a0, ref0)
out["rel_diff_culled"] = rel(a1, ref1)
out["ms_per_launch_8_chains"] = dict(dfma=ms_d8, fp32=ms_a8)
_, _, ms_d64 = run(64, 0, False)
_, _, ms_a64 = run(64, 1, False)
out["ms_per_launch_64_chains"] = dict(dfma=ms_d64, fp32=ms_a64)
print(json.dumps(out))
