import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import midaspom_b200 as mb
from gpu_util import make_engine, pdict
from stats_util import grid_marginals, grid_moments, ess
g = dict(np.load('tests/golden/golden.npz'))
obs = g['example_obs'].astype(np.int8)
grid, pe, pc, w = grid_marginals(g['post_default'])
print('exact e', grid_moments(grid, pe), 'c', grid_moments(grid, pc))
spec = dict(obs=obs, spacing=100.0, prior_occ=0.5)
for prec, name in ((mb.FP64, 'fp64'), (mb.FP32, 'fp32')):
    C, nsw, burn = 8, 12000, 1000
    with make_engine(spec, n_chains=C, precision=prec, seed=99, max_draws=nsw) as eng:
        eng.set_params([pdict(alpha=1/400)] * C)
        eng.init_chains(mb.engine.sampler_config(n_adapt=500, n_c_steps=2), disperse=True)
        eng.sweep(nsw)
        d = eng.get_draws()[burn:]
    for col, nm in ((0, 'e'), (1, 'c')):
        x = d[:, :, col]
        print(name, nm, 'mean', x.mean(), 'sd', x.std(), 'ess', sum(ess(x[:, i]) for i in range(C)), 'per-chain', np.round(x.mean(0), 3))
    print(name, 'ny1 mean', d[:, :, 6].mean(), 'nz1', d[:, :, 7].mean(), 'll', d[:, :, 5].mean())
