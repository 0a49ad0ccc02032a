"""Parameter recovery at full size: run the FP32 engine on a synthetic workload generated from known
parameters and compare the posterior (second half of the run) with the truth."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import midaspom_b200 as mb
from midaspom_b200 import synth, distributed as D
import bench

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
wl = synth.make_workload(name)
C = wl['chains_per_gpu']
eng = mb.Engine(wl['n'], wl['T'], C, precision=mb.FP32, seed=1000, detect=wl['detect'], max_draws=nsw)
eng.set_landscape_coords(wl['px'], wl['py'], wl['area']); eng.set_source_units(wl.get('src_unit'))
eng.set_observations(wl['obs']); eng.set_era(wl.get('era'))
eng.set_params([bench.start_params(wl)] * C)
kw = bench.sampler_kwargs(wl); kw['n_adapt'] = nsw // 4
eng.init_chains(mb.engine.sampler_config(**kw), disperse=(len(sys.argv) > 3 and sys.argv[3] == "disperse"))
t0 = time.time(); eng.sweep(nsw); dt = time.time() - t0
d = eng.get_draws()
half = d[nsw // 2:]
summ = D.posterior_summary(half[:, :, [0, 1, 2, 3, 4, 8, 9, 10]], fields=('e', 'c', 'alpha', 'b', 'p', 'K', 'Ksrc', 'dsrc'))
truth = wl['truth']
out = dict(workload=name, sweeps=nsw, seconds=dt, chain_iters_per_s=C * nsw / dt, truth=truth, posterior=summ,
           z_score={k: (summ[k]['mean'] - truth[k]) / summ[k]['sd'] for k in summ}, ess_per_s={k: v['ess'] / (dt / 2) for k, v in summ.items()},
           loglik_last=float(d[-1, :, 5].mean()), scales=eng.get_scales()[0].tolist())
print(json.dumps(out, indent=1))
