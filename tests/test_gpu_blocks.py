"""Block grid of the y scan (mp_set_scan_blocks): the scan visits the patches colour by colour, block by block, and the
FP32 engine scans the blocks of one colour concurrently.  No reference counterpart (the reference has no sampler): the
order is pinned against the oracle's spom_scan_order, the FP64 engine against the CPU twin draw by draw, and the
concurrent FP32 block scan against the FP64 engine flip by flip."""
import numpy as np
import pytest

import midaspom_b200 as mb
import oracle_lib as O
from gpu_util import make_engine, oparams, pdict, random_landscape

pytestmark = pytest.mark.gpu


def rel_close(got, want, rtol, floor=1e-300):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    ok = np.abs(got - want) <= rtol * np.maximum(np.abs(want), floor)
    assert ok.all(), (np.abs(got - want) / np.maximum(np.abs(want), floor)).max()


def model_with_blocks(spec, blocks):
    return O.Model(spec["obs"], geom=O.GEOM_COORDS, px=spec["px"], py=spec["py"], area=spec.get("area"),
                   prior_occ=spec.get("prior_occ", 0.5), detect=spec.get("detect", 0), era=spec.get("era"), scan_blocks=blocks)


@pytest.mark.parametrize("blocks", [(4, 3, 2), (7, 7, 3), (1, 5, 2), (2, 2, 2)])
def test_block_scan_order_equals_the_oracle_rule(blocks):
    rng = np.random.default_rng(5)
    spec, z, y = random_landscape(rng, 5003, 3, O.GEOM_COORDS)
    spec["px"][7] = spec["px"][21]; spec["py"][7] = spec["py"][21]      # a tie
    m = model_with_blocks(spec, blocks)
    with make_engine(spec, n_chains=1, precision=mb.FP32) as eng:
        plain = eng.scan_order()
        eng.set_scan_blocks(blocks[0], blocks[1], blocks[2], 10.0)
        got = eng.scan_order()
        assert (got == O.scan_order(m)).all()
        assert sorted(got.tolist()) == list(range(5003)) and (got != plain).any()
        eng.set_scan_blocks(1, 1, 1, 0.0)                                # off again: the plain Morton order
        assert (eng.scan_order() == plain).all()


def test_scan_blocks_rejects_grids_whose_same_colour_cells_touch_halos():
    rng = np.random.default_rng(6)
    spec, z, y = random_landscape(rng, 400, 3, O.GEOM_COORDS, side=10000.0)
    with make_engine(spec, n_chains=1, precision=mb.FP32) as eng:
        with pytest.raises(mb.MpError):
            eng.set_scan_blocks(10, 10, 2, 600.0)                        # cells 1 km wide, same colour 1 km apart < 2 x 600 m
        eng.set_scan_blocks(10, 10, 3, 600.0)                            # 2 km apart: fine
    lin = dict(obs=spec["obs"], geom=O.GEOM_LINEAR, spacing=100.0)
    with make_engine(lin, n_chains=1, precision=mb.FP32) as eng:
        with pytest.raises(mb.MpError):
            eng.set_scan_blocks(2, 2, 2, 1.0)                            # needs coordinates


def test_fp64_engine_with_blocks_follows_the_cpu_twin():
    """Same Philox counters, same (colour, block, Morton) order: identical latent states after 10 sweeps."""
    rng = np.random.default_rng(91)
    n, T, C, blocks = 120, 6, 3, (3, 2, 2)
    spec, z, _ = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.5, miss=0.08)
    m = model_with_blocks(spec, blocks)
    kw = dict(sample_alpha=1, sample_b=1, alpha_min=1e-4, alpha_max=5e-2, c_max=1.0, n_adapt=10, n_c_steps=2)
    par0 = pdict(e=0.4, c=0.03, alpha=1 / 400, b=0.5)
    ch = O.Chains(m, O.sampler_cfg(**kw), C, seed=4321, par0=oparams(par0), disperse=True)
    nsw = 10
    want = ch.run(nsw)
    with make_engine(spec, n_chains=C, seed=4321, max_draws=nsw) as eng:
        eng.set_scan_blocks(*blocks, 50.0)
        eng.set_params([par0] * C)
        eng.init_chains(mb.engine.sampler_config(**kw), disperse=True)
        eng.sweep(nsw)
        got = eng.get_draws()
        zg, yg = eng.get_state()
    assert (zg == ch.z).all() and (yg == ch.y).all()
    rel_close(got[:, :, :5], want[:, :, :5], 1e-9)


ALPHA = 1 / 100


def wide_landscape(rng, nside, T, occ=0.6):
    """Jittered lattice, 100 m spacing, dispersal range 1/alpha = 100 m: no patch is isolated (the smallest S stays around
    1e-2), so the halo is about 35 dispersal ranges and a landscape of nside = 145 (21,025 patches, 14.5 km) holds a 4 x 4
    grid with 3 x 3 colours -- the smallest problem on which blocks of one colour exist at all."""
    n = nside * nside
    gx, gy = np.meshgrid(np.arange(nside), np.arange(nside))
    spec = dict(geom=O.GEOM_COORDS, px=(gx.ravel() * 100.0 + rng.uniform(-30, 30, n)), py=(gy.ravel() * 100.0 + rng.uniform(-30, 30, n)),
                area=rng.lognormal(0, 0.5, n))
    z = (rng.random((T, n)) < occ).astype(np.uint8)
    y = (z[:-1] & z[1:] & (rng.random((T - 1, n)) < 0.6)).astype(np.uint8)
    obs = z.astype(np.int8)
    obs[rng.random((T, n)) < 0.05] = -1
    spec["obs"] = obs
    return spec, z, y


def auto_grid(spec, z, y, par, k, C=1):
    with make_engine(spec, n_chains=C, precision=mb.FP32, seed=1) as eng:
        eng.set_params([par] * C)
        eng.set_state(np.stack([z] * C), np.stack([y] * C))
        S0 = eng.connectivity()
        return eng.set_scan_blocks_auto(spec["px"], spec["py"], par["alpha"], float(S0.min()), float(np.max(spec["area"] ** par["b"])), k=k)


@pytest.mark.parametrize("nside,T,C,k,min_blocks", [(145, 3, 1, 3, 16), (145, 3, 2, 4, 25), (145, 4, 1, 2, 4)])
def test_fp32_block_scan_agrees_with_fp64_path(nside, T, C, k, min_blocks):
    """Blocks of one colour scanned concurrently by separate clusters, each with its own patches and the halo around them as
    targets: after one sweep from the same state the FP32 engine agrees with the FP64 engine (whole-year scan in the same
    order) flip by flip up to FP32 ties, S agrees, and the incrementally updated S equals a fresh recomputation."""
    rng = np.random.default_rng(900 + nside + k)
    spec, z, y = wide_landscape(rng, nside, T)
    par = pdict(e=0.35, c=0.05, alpha=ALPHA, b=0.5)
    grid = auto_grid(spec, z, y, par, k)
    assert grid[0] * grid[1] >= min_blocks, grid
    sc = mb.engine.sampler_config(sample_e=0, sample_c=0, update_z=0, n_adapt=0)
    out = {}
    for prec in (mb.FP64, mb.FP32):
        with make_engine(spec, n_chains=C, precision=prec, seed=11, max_draws=1) as eng:
            eng.set_scan_blocks(*grid)
            eng.set_params([par] * C)
            eng.set_state(np.stack([z] * C), np.stack([y] * C))
            eng.set_sampler(sc)
            eng.connectivity(fetch=False)
            eng.work_counters(reset=True)
            eng.sweep(1)
            zz, yy = eng.get_state()
            S_inc = eng.get_connectivity()
            work = eng.work_counters()
            geo = eng.scan_geometry()
            S_new = eng.connectivity()
            out[prec] = (yy, S_inc, S_new, work, geo)
    y64, S64, _, _, _ = out[mb.FP64]
    y32, S32, S32new, work, geo = out[mb.FP32]
    assert geo["blocks"], geo                                         # the block launches ran
    assert work["scan_blocks"] == C * (T - 1) * grid[0] * grid[1], (work["scan_blocks"], grid)   # every year of every chain by blocks
    cand = ((z[:-1] & z[1:]) == 1).sum() * C
    assert (y64 != y32).sum() <= max(2, 2e-3 * cand), (y64 != y32).sum()
    assert (y32 != np.stack([y] * C)).sum() > 0.05 * cand           # the scan moved
    assert ((y32 <= z[None, :-1]) & (y32 <= z[None, 1:])).all()
    rel_close(S32, S32new, 1e-6, floor=1e-9)
    same = (y64 == y32).all(axis=2)
    rel_close(S32[same], S64[same], 1e-5, floor=1e-9)


def test_block_scan_falls_back_to_whole_years_when_the_halo_is_too_small():
    """halo = 0: no year passes the validity check, every year is scanned as a whole in the block order -- and gives the
    same states as the concurrent block scan with a sufficient halo (the two are the same Gibbs scan)."""
    rng = np.random.default_rng(17)
    nside, T, C = 145, 3, 1
    spec, z, y = wide_landscape(rng, nside, T)
    par = pdict(e=0.35, c=0.05, alpha=ALPHA, b=0.5)
    sc = mb.engine.sampler_config(sample_e=0, sample_c=0, update_z=0, n_adapt=0)
    grid = auto_grid(spec, z, y, par, 3)
    assert grid[0] * grid[1] >= 16, grid
    res = []
    for halo in (grid[3], 0.0):
        with make_engine(spec, n_chains=C, precision=mb.FP32, seed=3, max_draws=2) as eng:
            eng.set_scan_blocks(grid[0], grid[1], grid[2], halo)
            eng.set_params([par] * C)
            eng.set_state(np.stack([z] * C), np.stack([y] * C))
            eng.set_sampler(sc)
            eng.connectivity(fetch=False)
            eng.work_counters(reset=True)
            eng.sweep(2)
            res.append((eng.get_state()[1], eng.get_connectivity(), eng.work_counters()["scan_blocks"]))
    assert res[0][2] > 0 and res[1][2] == 0
    cand = ((z[:-1] & z[1:]) == 1).sum() * C
    assert (res[0][0] != res[1][0]).sum() <= max(2, 1e-3 * cand), (res[0][0] != res[1][0]).sum()


def test_block_scan_with_a_marginal_halo_mixes_block_and_whole_year_scans():
    """A halo sized for the landscape's typical S but not for its smallest: the years whose smallest S is too small for it are
    scanned as a whole, the others by blocks, inside the same sweeps -- and the run agrees with the all-whole-year run (halo 0)
    up to FP32 ties, with S consistent with a recomputation."""
    rng = np.random.default_rng(23)
    nside, T, C = 145, 7, 1
    spec, z, y = wide_landscape(rng, nside, T)
    par = pdict(e=0.35, c=0.05, alpha=ALPHA, b=0.5)
    sc = mb.engine.sampler_config(sample_e=0, sample_c=0, update_z=0, n_adapt=0)
    with make_engine(spec, n_chains=C, precision=mb.FP32, seed=3) as eng:
        eng.set_params([par] * C)
        eng.set_state(np.stack([z] * C), np.stack([y] * C))
        S0 = eng.connectivity()[0]
    smin = np.sort(S0.min(axis=1))                                      # smallest S of every year
    apow = float(np.max(spec["area"] ** par["b"]))
    # the halo that is just enough for the year with the median smallest S
    halo = (36.0 * np.log(2.0) + np.log(apow / smin[len(smin) // 2])) / ALPHA * 1.0001
    ext = float(np.ptp(spec["px"]))
    nx = int(np.floor(2 * ext / (2.0 * halo) * 0.999))                  # k = 3
    assert nx >= 3, (nx, halo)
    res = []
    for h in (halo, 0.0):
        with make_engine(spec, n_chains=C, precision=mb.FP32, seed=3, max_draws=2) as eng:
            eng.set_scan_blocks(nx, nx, 3, h)
            eng.set_params([par] * C)
            eng.set_state(np.stack([z] * C), np.stack([y] * C))
            eng.set_sampler(sc)
            eng.connectivity(fetch=False)
            eng.work_counters(reset=True)
            eng.sweep(2)
            S_inc = eng.get_connectivity()
            res.append((eng.get_state()[1], S_inc, eng.work_counters()["scan_blocks"], eng.connectivity()))
    full = 2 * C * (T - 1) * nx * nx
    assert 0 < res[0][2] < full, (res[0][2], full)                      # some years by blocks, some as a whole
    assert res[1][2] == 0
    cand = ((z[:-1] & z[1:]) == 1).sum() * C
    assert (res[0][0] != res[1][0]).sum() <= max(2, 1e-3 * cand), (res[0][0] != res[1][0]).sum()
    rel_close(res[0][1], res[0][3], 1e-6, floor=1e-9)


def test_sharded_chain_with_blocks_equals_single_engine():
    """One chain over 3 emulated ranks (connectivity by target patches, scan by years) with the block grid on: the
    per-year block decisions depend on the year's own data only, so the sharded run equals the single engine bit for bit."""
    import torch
    from midaspom_b200 import distributed as D
    rng = np.random.default_rng(33)
    nside, T, C, W = 145, 4, 1, 3
    spec, z, y = wide_landscape(rng, nside, T)
    par = pdict(e=0.4, c=0.05, alpha=ALPHA, b=0.5)
    kw = dict(sample_alpha=1, sample_b=1, c_max=0.5, alpha_min=1e-3, alpha_max=1e-1, n_adapt=4)
    nsw = 4
    grid = auto_grid(spec, z, y, par, 3)

    def fresh():
        eng = make_engine(spec, n_chains=C, precision=mb.FP32, seed=17, max_draws=nsw)
        eng.set_scan_blocks(*grid)
        eng.set_params([par] * C)
        eng.init_chains(mb.engine.sampler_config(**kw), disperse=False)
        return eng

    ref = fresh()
    ref.work_counters(reset=True)
    ref.sweep(nsw)
    assert ref.work_counters()["scan_blocks"] > 0
    want = (ref.get_draws(), ref.get_state(), ref.get_connectivity())
    ref.close()
    engs = [fresh() for _ in range(W)]
    chains = [D.ShardedChain(e, r, W, torch.device("cuda", 0), reduce_fn=lambda t: None) for r, e in enumerate(engs)]
    D.sweep_emulated_ranks(chains, nsw)
    for e in engs:
        got = (e.get_draws(), e.get_state(), e.get_connectivity())
        assert (got[0] == want[0]).all()
        assert (got[1][0] == want[1][0]).all() and (got[1][1] == want[1][1]).all()
        assert (got[2] == want[2]).all()
        e.close()
