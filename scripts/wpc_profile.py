"""Phase split of the windowed y scan (build with MP_NVCC_EXTRA=-DMP_WPC_PROFILE).  usage: python scripts/wpc_profile.py [chains]"""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import midaspom_b200 as mb
from midaspom_b200 import synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = synth.make_workload("cfg3")
t = wl["truth"]
eng = mb.Engine(wl["n"], wl["T"], C, precision=mb.FP32, seed=1000, max_draws=16)
eng.set_landscape_coords(wl["px"], wl["py"], wl["area"]); eng.set_source_units(None)
eng.set_observations(wl["obs"])
eng.set_params([dict(e=0.5, c=t["c"], alpha=t["alpha"], b=t["b"])] * C)
eng.init_chains(mb.engine.sampler_config(sample_alpha=1, sample_b=1, c_max=20 * t["c"], alpha_min=t["alpha"] / 5, alpha_max=t["alpha"] * 5, n_adapt=50), disperse=False)
eng.sweep(3); eng.synchronize()
eng.set_timing(True); eng.get_timing(reset=True); eng.work_counters(reset=True)
nsw = 5
eng.sweep(nsw); eng.synchronize()
ms, _ = eng.get_timing(reset=True)
w = eng.work_counters(); d = eng.debug_counters()
ntask = C * (wl["T"] - 1)
cyc = np.array(d[:5], dtype=float) / (nsw * ntask); cnt = np.array(d[5:10], dtype=float) / (nsw * ntask)
print(json.dumps(dict(chains=C, scan_ms=ms["sweep_y"] / nsw, cycles_per_task=dict(zip(("eval", "decide", "second", "commit", "other"), cyc.round(0).tolist())),
                      per_task=dict(zip(("rounds", "second", "careful", "commits", "retired"), cnt.round(1).tolist())),
                      cycles_per_round=float((cyc.sum() / max(cnt[0], 1)).round(0)), retired_per_round=float((cnt[4] / max(cnt[0], 1)).round(2)),
                      exec_groups_per_cand=w["scan_exec"] / max(d[9], 1), commit_groups_per_commit=w["scan_commit"] / max(d[8], 1), geo=eng.scan_geometry())))
eng.close()
