"""Time the y scan of one large chain with and without the block grid (mp_set_scan_blocks).
usage: python scripts/blocks_micro.py WORKLOAD K [SWEEPS]      (K = 0: no grid)"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import midaspom_b200 as mb
from midaspom_b200 import synth

name, k = sys.argv[1], int(sys.argv[2])
nsw = int(sys.argv[3]) if len(sys.argv) > 3 else 4
wl = synth.make_workload(name)
n, T, C = wl["n"], wl["T"], wl["chains_per_gpu"]
t = wl["truth"]
eng = mb.Engine(n, T, C, precision=mb.FP32, seed=1000, max_draws=nsw + 8)
eng.set_landscape_coords(wl["px"], wl["py"], wl["area"]); eng.set_source_units(None)
eng.set_observations(wl["obs"])
eng.set_params([dict(e=0.5, c=t["c"], alpha=t["alpha"], b=t["b"])] * C)
eng.init_chains(mb.engine.sampler_config(sample_alpha=1, sample_b=1, c_max=20 * t["c"], alpha_min=t["alpha"] / 5, alpha_max=t["alpha"] * 5, n_adapt=50), disperse=False)
S0 = eng.get_connectivity()
grid = None
if k:
    grid = eng.set_scan_blocks_auto(wl["px"], wl["py"], t["alpha"], float(S0.min()), float(np.max(wl["area"] ** t["b"])), k=k)
eng.sweep(2); eng.synchronize()
eng.set_timing(True); eng.get_timing(reset=True); eng.work_counters(reset=True)
t0 = time.perf_counter()
eng.sweep(nsw); eng.synchronize()
dt = (time.perf_counter() - t0) / nsw
ms, launches = eng.get_timing(reset=True)
work = eng.work_counters()
d = eng.get_draws()
print(json.dumps(dict(workload=name, k=k, grid=grid, s_min=float(S0.min()), ms_per_sweep=dt * 1e3, kernel_ms={a: b / nsw for a, b in ms.items()},
                      launches={a: b / nsw for a, b in launches.items()}, scan=eng.scan_geometry(), blocks_per_sweep=work["scan_blocks"] / nsw,
                      trips=work["scan_trips"] / nsw, loglik=float(d[-1, 0, 5]), n_y1=float(d[-1, 0, 6]))))
eng.close()
