#!/bin/bash
# debug: swap in the -DMP_DEBUG_CULL build and print the fraction of slots evaluated per flip
cp midaspom_b200/lib/libdbg.so midaspom_b200/lib/libmidaspom_cuda.so
touch midaspom_b200/lib/libmidaspom_cuda.so
for wl in cfg3 cfg5t; do
MP_FAST_CULL=1 timeout 600 python - $wl <<'PY'
import sys, ctypes as C, numpy as np
import bench
import midaspom_b200 as mb
from midaspom_b200 import synth
wl = synth.make_workload(sys.argv[1])
n, T, cpg = wl["n"], wl["T"], wl["chains_per_gpu"]
eng = mb.Engine(n, T, cpg, precision=mb.FP32, device=0, seed=1000, detect=wl["detect"], max_draws=64)
eng.set_landscape_coords(wl["px"], wl["py"], wl["area"]); eng.set_source_units(None)
eng.set_observations(wl["obs"]); eng.set_era(wl.get("era"))
eng.set_params([bench.start_params(wl)] * cpg)
eng.init_chains(mb.engine.sampler_config(**bench.sampler_kwargs(wl)), disperse=False)
eng.sweep(3)
L = mb.engine.load_library()
out = (C.c_ulonglong * 4)()
print("rc", L.mp_debug_cull_counters(out))
print(sys.argv[1], "warp-flips", out[0], "active", out[1], "total", out[2], "fraction", out[1] / max(out[2], 1))
PY
done
