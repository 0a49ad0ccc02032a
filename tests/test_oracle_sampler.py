"""The CPU twin of the CUDA sampler (oracle spom_sweep) targets the reference's posterior: on the
bundled example its (e, c) draws must agree with the exact 101x101 grid posterior that
MIDASPOM.out writes (golden post_default).  This validates the ALGORITHM; the GPU tests then
check the CUDA engine against this twin draw by draw."""
import numpy as np

import oracle_lib as O
from stats_util import grid_marginals, grid_moments, ess, ks_distance_thinned, rhat


def test_cpu_sampler_matches_exact_grid_posterior(golden, example_obs):
    m = O.Model(example_obs, spacing=100.0, prior_occ=0.5)
    cfg = O.sampler_cfg(n_adapt=500, n_e_steps=4, n_c_steps=2)
    ch = O.Chains(m, cfg, 4, seed=2026, par0=O.params(alpha=1 / 400), disperse=True)
    d = ch.run(16000)[1000:]
    grid, pe, pc, w = grid_marginals(golden["post_default"])
    for col, pm in ((0, pe), (1, pc)):
        mean, sd = grid_moments(grid, pm)
        x = d[:, :, col]
        n_eff = sum(ess(x[:, i]) for i in range(x.shape[1]))
        assert abs(x.mean() - mean) < 5 * sd / np.sqrt(n_eff) + 2e-3      # grid discretisation ~1e-3
        assert abs(x.std() - sd) < 0.06 * sd
        assert rhat(x) < 1.02
        thin = max(1, int(np.ceil(x.shape[0] * x.shape[1] / n_eff)) * 2)
        dist, n = ks_distance_thinned(x.T.ravel(), grid, pm, thin)
        assert dist < 1.63 / np.sqrt(n) + 0.01                              # KS 1% level + grid step
    exact_corr = ((w * np.outer(grid - grid_moments(grid, pe)[0], grid - grid_moments(grid, pc)[0])).sum()
                  / (grid_moments(grid, pe)[1] * grid_moments(grid, pc)[1]))
    assert abs(np.corrcoef(d[:, :, 0].ravel(), d[:, :, 1].ravel())[0, 1] - exact_corr) < 0.05


def test_sweep_keeps_state_feasible_and_S_consistent():
    """After sweeps: y <= z_t & z_t+1, observed cells untouched, and the incrementally updated S
    equals a from-scratch recomputation (rank-1 bookkeeping)."""
    rng = np.random.default_rng(11)
    n, T = 60, 6
    px, py = rng.uniform(0, 4000, n), rng.uniform(0, 4000, n)
    z0 = (rng.random(n) < 0.5).astype(np.uint8)
    m0 = O.Model(np.zeros((T, n), dtype=np.int8), geom=O.GEOM_COORDS, px=px, py=py)
    truth = O.params(e=0.3, c=0.08, alpha=1 / 600)
    ztrue = O.simulate(m0, truth, 5, 0, z0, T - 1)
    obs = ztrue.astype(np.int8)
    obs[1:][rng.random((T - 1, n)) < 0.1] = -1
    m = O.Model(obs, geom=O.GEOM_COORDS, px=px, py=py, detect=0)
    cfg = O.sampler_cfg(sample_alpha=1, alpha_min=1e-4, alpha_max=1e-2, c_max=2.0)
    ch = O.Chains(m, cfg, 2, seed=5, par0=O.params(e=0.3, c=0.08, alpha=1 / 600), disperse=False)
    ch.run(30)
    for c in range(2):
        z, y = ch.z[c], ch.y[c]
        assert ((y <= z[:-1]) & (y <= z[1:])).all()
        assert (z[obs == 1] == 1).all() and (z[obs == 0] == 0).all()
        par = ch.par[c]
        S = np.array([O.connectivity(m, par.alpha, par.b, y[t]) for t in range(T - 1)])
        # S is refreshed at the start of a sweep and then updated incrementally through the y scan
        np.testing.assert_allclose(ch.S[c], S, rtol=1e-10, atol=1e-14)
        assert np.isfinite(ch.draws[c, 5])


def test_scan_order_is_the_morton_curve_of_planar_coordinates():
    """The y scan visits planar landscapes along the Z-order curve (16 bits per axis over the bounding square, ties by
    index) and linear / dense ones in index order -- the rule the CUDA engine shares (mp_get_scan_order)."""
    rng = np.random.default_rng(5)
    obs = np.zeros((2, 4), dtype=np.int8)
    # unit square corners: Z order is (0,0), (1,0), (0,1), (1,1) -- x is the low bit
    m = O.Model(obs, geom=O.GEOM_COORDS, px=np.array([1.0, 0.0, 1.0, 0.0]), py=np.array([1.0, 1.0, 0.0, 0.0]))
    assert O.scan_order(m).tolist() == [3, 2, 1, 0]
    n = 777
    obs = np.zeros((2, n), dtype=np.int8)
    px, py = rng.uniform(0, 5000, n), rng.uniform(0, 3000, n)
    px[10] = px[500]; py[10] = py[500]                            # identical positions: index order decides
    order = O.scan_order(O.Model(obs, geom=O.GEOM_COORDS, px=px, py=py))
    assert sorted(order.tolist()) == list(range(n))
    assert list(order).index(10) + 1 == list(order).index(500)
    # independent restatement of the code: interleave the bits of the quantised coordinates
    span = max(px.max() - px.min(), py.max() - py.min())
    xi = np.minimum(65535.0, (px - px.min()) / span * 65535.0).astype(np.uint64)
    yi = np.minimum(65535.0, (py - py.min()) / span * 65535.0).astype(np.uint64)
    code = np.zeros(n, dtype=np.uint64)
    for bit in range(16):
        code |= ((xi >> np.uint64(bit)) & np.uint64(1)) << np.uint64(2 * bit)
        code |= ((yi >> np.uint64(bit)) & np.uint64(1)) << np.uint64(2 * bit + 1)
    assert (order == np.argsort(code, kind="stable")).all()
    # consecutive visits are spatial neighbours: far closer than two random patches
    step = np.hypot(np.diff(px[order]), np.diff(py[order]))
    assert np.median(step) < 0.15 * np.median(np.hypot(px - px[::-1], py - py[::-1]))
    assert O.scan_order(O.Model(obs, geom=O.GEOM_LINEAR, spacing=100.0)).tolist() == list(range(n))


def test_scan_order_with_a_block_grid_visits_colour_by_colour():
    """Block grid of the scan (spom_model.blk_nx, blk_ny, blk_k; engine: mp_set_scan_blocks): the order is a permutation that
    visits the cells colour by colour (colour = (bx mod k) + k (by mod k)), cell by cell inside a colour, and in plain Morton
    order inside a cell; a 1 x 1 grid is the plain Morton order."""
    rng = np.random.default_rng(3)
    n, nx, ny, k = 2000, 5, 4, 2
    px, py = rng.uniform(0, 1000.0, n), rng.uniform(0, 800.0, n)
    obs = np.zeros((2, n), dtype=np.int8)
    plain = O.scan_order(O.Model(obs, geom=O.GEOM_COORDS, px=px, py=py))
    assert (O.scan_order(O.Model(obs, geom=O.GEOM_COORDS, px=px, py=py, scan_blocks=(1, 1, 1))) == plain).all()
    order = O.scan_order(O.Model(obs, geom=O.GEOM_COORDS, px=px, py=py, scan_blocks=(nx, ny, k)))
    assert sorted(order.tolist()) == list(range(n))
    bx = np.minimum(((px - px.min()) / ((px.max() - px.min()) / nx)).astype(int), nx - 1)
    by = np.minimum(((py - py.min()) / ((py.max() - py.min()) / ny)).astype(int), ny - 1)
    key = ((bx % k) + k * (by % k)) * (nx * ny) + by * nx + bx            # (colour, cell)
    assert (np.diff(key[order]) >= 0).all()                                # colour-major, then cell
    rank = np.empty(n, dtype=int); rank[plain] = np.arange(n)              # position in the plain Morton order
    for g in np.unique(key):
        inside = order[key[order] == g]
        assert (np.diff(rank[inside]) > 0).all()                           # Morton order inside a cell
