"""GPU parity tests: the CUDA engine, called through the C ABI (libmidaspom_cuda.so), against the
CPU oracle on the same seeded inputs, the committed golden fixtures, and size-independent
properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): log-likelihood and connectivity within 1e-9 relative on
the FP64 path and 1e-5 on the FP32 path; latent-state indexing and bookkeeping bit-exact."""
import numpy as np
import pytest

import midaspom_b200 as mb
import oracle_lib as O
from gpu_util import make_engine, make_model, pdict, oparams, random_landscape
from stats_util import grid_marginals, grid_moments, ess, ks_distance_thinned, rhat

pytestmark = pytest.mark.gpu

RTOL = {mb.FP64: 1e-9, mb.FP32: 1e-5}
A = 1.0 / 400


def rel_close(got, want, rtol, floor=1e-300):
    got, want = np.asarray(got, float), np.asarray(want, float)
    both_inf = np.isneginf(got) & np.isneginf(want)
    ok = both_inf | (np.abs(got - want) <= rtol * np.maximum(np.abs(want), floor))
    assert ok.all(), f"max rel err {np.nanmax(np.abs(got - want) / np.maximum(np.abs(want), floor)):.3e} > {rtol}"


def oracle_S(m, par, y):
    return np.array([O.connectivity(m, par["alpha"], par["b"], y[t]) for t in range(y.shape[0])])


# ------------------------------------------------------------------ connectivity (a2)
@pytest.mark.parametrize("precision", [mb.FP64, mb.FP32])
def test_connectivity_bundled_example_every_state(golden, precision):
    """S for all 256 enumerated states of the bundled example == main_MIDASPOM.c:351-355."""
    piall = golden["cpp_piall"].astype(np.uint8)            # 256 x 8
    n = piall.shape[1]
    spec = dict(obs=np.zeros((2, n), dtype=np.int8), spacing=100.0)
    m = make_model(spec)
    C = piall.shape[0]
    with make_engine(spec, n_chains=C, precision=precision) as eng:
        eng.set_params([pdict(alpha=A)] * C)
        eng.set_state(np.zeros((C, 2, n), np.uint8), piall[:, None, :])
        S = eng.connectivity()
    want = np.array([O.connectivity(m, A, 0.0, piall[j]) for j in range(C)])
    rel_close(S[:, 0, :], want, 1e-14 if precision == mb.FP64 else 1e-5, floor=1e-12)
    assert (S[0] == 0).all()                                 # empty landscape: exactly zero


@pytest.mark.parametrize("precision", [mb.FP64, mb.FP32])
@pytest.mark.parametrize("geom", [O.GEOM_LINEAR, O.GEOM_COORDS, O.GEOM_DENSE])
def test_connectivity_random_landscapes(precision, geom):
    rng = np.random.default_rng(100 + geom)
    n, T, C = 517, 11, 3                                      # ragged: not a multiple of the 128-wide tile
    spec, z, y = random_landscape(rng, n, T, geom)
    m = make_model(spec)
    pars = [pdict(alpha=1 / 300, b=0.0), pdict(alpha=1 / 500, b=0.5), pdict(alpha=1 / 150, b=1.3)]
    ys = np.stack([y, np.roll(y, 1, axis=1), np.zeros_like(y)])
    with make_engine(spec, n_chains=C, precision=precision) as eng:
        eng.set_params(pars)
        eng.set_state(np.stack([z] * C), ys)
        S = eng.connectivity()
    for c in range(C):
        rel_close(S[c], oracle_S(m, pars[c], ys[c]), RTOL[precision], floor=1e-6)
    assert (S[2] == 0).all()


def test_connectivity_more_than_32_years():
    rng = np.random.default_rng(5)
    spec, z, y = random_landscape(rng, 70, 41, O.GEOM_COORDS)
    m = make_model(spec)
    par = pdict(alpha=1 / 400, b=0.4)
    with make_engine(spec, precision=mb.FP64) as eng:
        eng.set_params([par]); eng.set_state(z[None], y[None])
        rel_close(eng.connectivity()[0], oracle_S(m, par, y), 1e-12, floor=1e-9)


# ------------------------------------------------------------------ log-likelihood (a3, a4, a8)
CASES = [
    ("base", dict(), pdict(e=0.31, c=0.012, alpha=1 / 400)),
    ("areas", dict(), pdict(e=0.55, c=0.02, alpha=1 / 250, b=0.8)),
    ("dieoff", dict(era=True), pdict(e=0.4, c=0.01, alpha=1 / 400, K=2.5)),
    ("loss", dict(era=True), pdict(e=0.4, c=0.01, alpha=1 / 400, K=1.0, Ksrc=3.0, dsrc=35.0)),
    ("future", dict(era=True), pdict(e=0.9, c=0.015, alpha=1 / 300, K=1.7, Ksrc=0.8, dsrc=50.0, b=0.3)),
    ("detect", dict(detect=1), pdict(e=0.3, c=0.02, alpha=1 / 400, p=0.8)),
    ("clampE", dict(), pdict(e=1.4, c=0.02, alpha=1 / 400)),
]


@pytest.mark.parametrize("precision", [mb.FP64, mb.FP32])
@pytest.mark.parametrize("geom", [O.GEOM_LINEAR, O.GEOM_COORDS])
@pytest.mark.parametrize("name,extra,par", CASES, ids=[c[0] for c in CASES])
def test_loglik_parts_vs_oracle(precision, geom, name, extra, par):
    rng = np.random.default_rng(sum(map(ord, name)) + geom)
    n, T = 300, 9
    spec, z, y = random_landscape(rng, n, T, geom, miss=0.05 if name != "detect" else 0.0)
    if extra.get("era"):
        spec["era"] = (np.arange(T - 1) < 4).astype(np.uint8)
    if extra.get("detect"):
        spec["detect"] = 1
        obs = z.astype(np.int8)
        obs[(z == 1) & (rng.random(z.shape) < 0.2)] = 0      # missed detections
        spec["obs"] = obs
    if name == "clampE":
        y[:] = 0                                              # E = 1: every occupied patch goes extinct
    m = make_model(spec)
    want, parts = O.loglik(m, oparams(par), z, y)
    with make_engine(spec, precision=precision) as eng:
        ll, gparts = eng.loglik_host([par], z[None], y[None])
    rel_close(gparts[0], parts, RTOL[precision], floor=1.0)
    rel_close(ll[0], want, RTOL[precision], floor=1.0)


def test_loglik_impossible_states_are_minus_infinity():
    """Where the reference returns probability 0 (compPePc:34-37,46-47) the engine returns -inf."""
    n, T = 16, 3
    obs = np.zeros((T, n), np.int8); obs[:, :8] = 1
    spec = dict(obs=obs, spacing=100.0)
    z = (obs == 1).astype(np.uint8)
    y_ok = (z[:-1] & z[1:]).astype(np.uint8)
    par = pdict(e=0.3, c=0.5, alpha=A)
    bad_y0 = y_ok.copy(); bad_y0[0, 12] = 1                   # y=1 where z_t=0
    z_bad = z.copy(); z_bad[1, 3] = 0                         # contradicts obs=1 under perfect detection
    y_none = np.zeros_like(y_ok)                              # nobody survives but patches occupied next year: C=0
    y_forced = y_ok.copy(); z2 = z.copy(); z2[2, 0] = 0       # y=1 but z'=0
    with make_engine(spec, n_chains=4) as eng:
        ll, parts = eng.loglik_host([par] * 4, np.stack([z, z_bad, z, z2]), np.stack([bad_y0, y_ok, y_none, y_forced]))
    assert np.isneginf(ll).all()
    m = make_model(spec)
    for zz, yy in ((z, bad_y0), (z_bad, y_ok), (z, y_none), (z2, y_forced)):
        assert np.isneginf(O.loglik(m, oparams(par), zz, yy)[0])
    zc = z.copy()                                             # clamp C=1 with z'=0 => log(1-1)
    spec2 = dict(obs=zc.astype(np.int8), spacing=100.0)
    with make_engine(spec2) as eng:
        ll, _ = eng.loglik_host([pdict(e=0.3, c=50.0, alpha=A)], zc[None], (zc[:-1] & zc[1:])[None])
        assert np.isneginf(ll[0])


def test_loglik_bundled_example_complete_data_terms(golden, example_obs):
    """Complete-data log-likelihood on the bundled example for every completion of its 3 missing
    cells, y = all survivors: engine == oracle (which is pinned to compPePc)."""
    T, n = example_obs.shape
    miss = np.argwhere(example_obs == -1)
    zs, ys = [], []
    for mask in range(2 ** len(miss)):
        z = (example_obs == 1).astype(np.uint8)
        for b, (t, k) in enumerate(miss):
            z[t, k] = (mask >> b) & 1
        zs.append(z); ys.append(z[:-1] & z[1:])
    spec = dict(obs=example_obs, spacing=100.0, prior_occ=0.5)
    m = make_model(spec)
    par = pdict(e=0.71, c=0.52, alpha=A)
    with make_engine(spec, n_chains=len(zs)) as eng:
        ll, parts = eng.loglik_host([par] * len(zs), np.stack(zs), np.stack(ys))
    want = [O.loglik(m, oparams(par), z, y)[0] for z, y in zip(zs, ys)]
    rel_close(ll, want, 1e-12, floor=1.0)


# ------------------------------------------------------------------ exact small-n engine (a5-a7): the reference's own output
def test_exact_engine_reproduces_reference_posterior_tables(golden, example_obs):
    """mp_exact_posterior == MIDASPOM.out on the bundled example: the 101x101 table of
    run_examples.sh:8, the -s 11 table with its zero-likelihood edges, the -s 3 -l .3 -u .7 grid,
    the 'Total log-likelihood' line, and the state bookkeeping the program prints."""
    for key, nstep, lo, hi in (("post_default", 101, 0.0, 1.0), ("post_s11", 11, 0.0, 1.0), ("post_s3", 3, 0.3, 0.7)):
        ll, ltot, info = mb.exact_posterior(example_obs, a=A, d=100.0, prior_occ=0.5, nstep=nstep, ecmin=lo, ecmax=hi)
        tab = golden[key]
        assert info == dict(nvar=8, nstates=256, nextid=10, max_states_per_year=2)       # main_MIDASPOM.c:300-304 echo
        assert abs(ltot - float(golden["ltot_" + key.split("_")[1]])) < 6e-6             # printed with %.5lf
        dens = np.exp(ll - ltot)
        assert (dens[tab == 0] < 5.1e-21).all()                                          # printed as 0.00000000000000000000
        np.testing.assert_allclose(dens, tab, rtol=1e-9, atol=5.1e-21)                   # table printed with %.20lf
    ll3, _, _ = mb.exact_posterior(example_obs, a=A, d=100.0, nstep=3, ecmin=0.3, ecmax=0.7)
    np.testing.assert_allclose(ll3, golden["loglik_s3_survey"], rtol=0, atol=2e-12)


def test_exact_engine_on_wide_inputs_matches_the_reference():
    """More than 32 distinct observation-compatible rows (P tiled over the short list) and a year with 6 missing cells
    (64 completions): the posterior tables the reference binary printed for these inputs (tests/golden/make_golden_wide.py)."""
    from pathlib import Path
    g = dict(np.load(Path(__file__).resolve().parent / "golden" / "golden_wide.npz"))
    for tag, nstep, min_next in (("a", 21, 70), ("b", 11, 34)):
        obs, tab, want_ltot = g[f"wide_{tag}_obs"].astype(np.int8), g[f"wide_{tag}_post"], float(g[f"wide_{tag}_ltot"])
        ll, ltot, info = mb.exact_posterior(obs, a=A, d=100.0, prior_occ=0.5, nstep=nstep, ecmin=0.0, ecmax=1.0)
        assert info["nextid"] >= min_next, info
        assert abs(ltot - want_ltot) < 6e-6                                              # printed with %.5lf
        np.testing.assert_allclose(np.exp(ll - ltot), tab, rtol=1e-9, atol=5.1e-21)
    # the oracle's forward recursion agrees too, at full double precision
    m = O.Model(g["wide_a_obs"].astype(np.int8), spacing=100.0, prior_occ=0.5)
    ll, _, _ = mb.exact_posterior(g["wide_a_obs"].astype(np.int8), a=A, d=100.0, nstep=3, ecmin=0.2, ecmax=0.8)
    want = np.array([[O.marginal_loglik(m, O.params(e=e, c=c, alpha=A)) for c in (0.2, 0.5, 0.8)] for e in (0.2, 0.5, 0.8)])
    rel_close(ll, want, 1e-10, floor=1.0)


def test_exact_variant_beyond_shared_memory(monkeypatch, example_obs):
    """14-16 patches keep the two state vectors in global memory: the same code path forced on the bundled example must
    reproduce the shared-memory result bit for bit, and a 14-patch row must run and give a likelihood in (0, 1]."""
    small = mb.exact_variant("dieoff", example_obs[0], 0.71, 0.52, ts=6, tdis=4, a=A, d=100.0, nstepK=9)
    monkeypatch.setenv("MP_EXACT_GLOBAL", "1")
    forced = mb.exact_variant("dieoff", example_obs[0], 0.71, 0.52, ts=6, tdis=4, a=A, d=100.0, nstepK=9)
    assert (small == forced).all()
    monkeypatch.delenv("MP_EXACT_GLOBAL")
    row = np.array([1, 0, 1, 1, -1, 0, 1, 0, 0, 1, 1, 0, -1, 1], dtype=np.int8)
    big = mb.exact_variant("loss", row, 0.5, 0.3, ts=3, tdis=2, a=A, d=100.0, nstepK=3, nstepd=2)
    assert big.shape == (3, 2) and np.isfinite(big).all() and (big > 0).all() and (big <= 1).all()
    with pytest.raises(mb.MpError):
        mb.exact_variant("dieoff", np.ones(17, dtype=np.int8), 0.5, 0.3, ts=1, tdis=1)


def test_exact_variants_reproduce_dieoff_and_loss_outputs(golden, example_obs):
    """run_examples.sh:11,14 -- `MIDASPOM_dieoff.out -a 10 -e 0.71 -c 0.52 -m 400 -d 100 -s 151` (151 values) and
    `MIDASPOM_loss.out ... -s 7 -v 3` (7 x 3): vector propagation here vs matrix powers there."""
    got = mb.exact_variant("dieoff", example_obs[0], 0.71, 0.52, ts=20, tdis=10, a=A, d=100.0, nstepK=151)
    np.testing.assert_allclose(got, golden["dieoff_s151"], rtol=1e-9, atol=5.1e-21)   # reference prints %.20lf
    got = mb.exact_variant("loss", example_obs[0], 0.71, 0.52, ts=20, tdis=10, a=A, d=100.0, nstepK=7, nstepd=3)
    np.testing.assert_allclose(got, golden["loss_s7v3"], rtol=1e-9, atol=5.1e-21)


def test_exact_engine_matches_oracle_marginal_on_random_small_landscapes():
    """Random 6-10 patch histories with missing cells (also in year 0, where the reference itself
    reads uninitialised memory): exact engine == oracle's forward recursion."""
    rng = np.random.default_rng(77)
    for n, T, miss0 in ((6, 5, False), (9, 4, True), (10, 6, True), (7, 3, False)):
        z = (rng.random((T, n)) < 0.55).astype(np.int8)
        z[:, rng.integers(n)] = 0                                  # a never-occupied patch: dropped from the enumeration (:203-204)
        obs = z.copy()
        obs[1:][rng.random((T - 1, n)) < 0.12] = -1
        if miss0:
            obs[0, rng.choice(n, 2, replace=False)] = -1
        m = O.Model(obs, spacing=150.0, prior_occ=0.25)
        ll, ltot, info = mb.exact_posterior(obs, a=1 / 300, d=150.0, prior_occ=0.25, nstep=5, ecmin=0.1, ecmax=0.9)
        grid = [0.1, 0.3, 0.5, 0.7, 0.9]
        want = np.array([[O.marginal_loglik(m, O.params(e=e, c=c, alpha=1 / 300)) for c in grid] for e in grid])
        rel_close(ll, want, 1e-10, floor=1.0)


# ------------------------------------------------------------------ rank-1 flips
@pytest.mark.parametrize("precision", [mb.FP64, mb.FP32])
@pytest.mark.parametrize("geom", [O.GEOM_LINEAR, O.GEOM_COORDS, O.GEOM_DENSE])
def test_flip_delta_vs_oracle(precision, geom):
    rng = np.random.default_rng(40 + geom)
    n, T = 200, 6
    spec, z, y = random_landscape(rng, n, T, geom)
    spec["era"] = np.array([1, 1, 0, 0, 0], dtype=np.uint8)
    m = make_model(spec)
    par = pdict(e=0.35, c=0.015, alpha=1 / 450, b=0.6, K=1.8, Ksrc=0.5, dsrc=30.0)
    cand = np.argwhere((z[:-1] & z[1:]) == 1)
    pick = cand[rng.choice(len(cand), 12, replace=False)]
    with make_engine(spec, precision=precision) as eng:
        eng.set_params([par]); eng.set_state(z[None], y[None])
        S = eng.connectivity()[0]
        got = [eng.flip_delta(0, int(t), int(k)) for t, k in pick]
    want = [O.flip_delta_bruteforce(m, oparams(par), z, y, int(t), int(k)) for t, k in pick]
    tol = 1e-9 if precision == mb.FP64 else 2e-4               # sum of ~200 FP32 log differences
    for g, w in zip(got, want):
        assert abs(g - w) <= tol * max(1.0, abs(w))


# ------------------------------------------------------------------ sampler: draw-by-draw twin
@pytest.mark.parametrize("geom,detect,sample_ab", [(O.GEOM_LINEAR, 0, 0), (O.GEOM_COORDS, 0, 1), (O.GEOM_COORDS, 1, 1)])
def test_fp64_sweeps_follow_the_cpu_twin(geom, detect, sample_ab):
    """Same Philox counters, same scan order: after k sweeps the FP64 engine and the oracle twin
    hold IDENTICAL latent states (bit-exact bookkeeping) and parameters equal to rounding."""
    rng = np.random.default_rng(77 + geom + detect)
    n, T, C = 90, 7, 3
    spec, z, _ = random_landscape(rng, n, T, geom, occ=0.5, miss=0.08, areas=bool(sample_ab))
    if detect:
        spec["detect"] = 1
    m = make_model(spec)
    kw = dict(sample_alpha=sample_ab, sample_b=sample_ab, sample_p=detect, alpha_min=1e-4, alpha_max=5e-2,
              c_max=1.0, n_adapt=10, n_c_steps=2)
    par0 = pdict(e=0.4, c=0.03, alpha=1 / 400, b=0.5 if sample_ab else 0.0, p=0.8 if detect else 1.0)
    seed = 1234
    ch = O.Chains(m, O.sampler_cfg(**kw), C, seed=seed, par0=oparams(par0), disperse=True)
    nsw = 12
    want = ch.run(nsw)
    with make_engine(spec, n_chains=C, seed=seed, max_draws=nsw) as eng:
        eng.set_params([par0] * C)
        eng.init_chains(mb.engine.sampler_config(**kw), disperse=True)
        eng.sweep(nsw)
        got = eng.get_draws()
        zg, yg = eng.get_state()
        Sg = eng.get_connectivity()
    assert (zg == ch.z).all() and (yg == ch.y).all()
    assert (got[:, :, 6:] == want[:, :, 6:]).all()               # counts of y=1 and z=1, every sweep
    rel_close(got[:, :, :5], want[:, :, :5], 1e-9)
    rel_close(got[:, :, 5], want[:, :, 5], 1e-9, floor=1.0)
    rel_close(Sg, ch.S, 1e-9, floor=1e-9)


@pytest.mark.parametrize("geom,n,C,T,variant", [(O.GEOM_COORDS, 700, 3, 6, None), (O.GEOM_LINEAR, 2100, 2, 4, None),
                                                 (O.GEOM_DENSE, 300, 5, 5, None), (O.GEOM_COORDS, 1500, 19, 9, None),
                                                 (O.GEOM_COORDS, 900, 4, 7, "dieoff"), (O.GEOM_LINEAR, 600, 3, 6, "loss"),
                                                 (O.GEOM_COORDS, 300, 2, 41, None),
                                                 # above 3,000 patches the spatially culled scan (k_sweep_y_cull) and the culled k_conn run
                                                 (O.GEOM_COORDS, 6000, 2, 4, None), (O.GEOM_LINEAR, 5200, 2, 3, None),
                                                 (O.GEOM_COORDS, 4100, 3, 5, "dieoff"),
                                                 # beyond one CTA's shared memory in both precisions: FP32 cluster of 8 x 256 threads with 4
                                                 # candidates per exchange, FP64 scan with its per-task state in global scratch
                                                 (O.GEOM_COORDS, 40000, 1, 3, None), (O.GEOM_COORDS, 20000, 2, 3, None)])
def test_fp32_fast_sweep_agrees_with_fp64_path(geom, n, C, T, variant):
    """The FP32 throughput sweep (cluster-split, two-float S, product-of-factors logs) draws from the
    same thresholds as the FP64 path: after one sweep from the same state the latent states agree
    except where a log-odds lies within FP32 rounding of its threshold, S agrees to 1e-5, and the
    incrementally updated S equals a fresh recomputation (rank-1 bookkeeping)."""
    rng = np.random.default_rng(500 + n)
    spec, z, y = random_landscape(rng, n, T, geom, occ=0.5, miss=0.05)
    par = pdict(e=0.35, c=0.02 if geom != O.GEOM_LINEAR else 0.2, alpha=1 / 400, b=0.5)
    if variant:                                                   # pre-event years with the dieoff / loss terms
        spec["era"] = (np.arange(T - 1) < (T - 1) // 2).astype(np.uint8)
        par.update(dict(K=2.2) if variant == "dieoff" else dict(K=1.0, Ksrc=4.0, dsrc=3.0))
    sc = mb.engine.sampler_config(sample_e=0, sample_c=0, update_z=0, n_adapt=0)
    out = {}
    for prec in (mb.FP64, mb.FP32):
        with make_engine(spec, n_chains=C, precision=prec, seed=11, max_draws=1) as eng:
            eng.set_params([par] * C)
            eng.set_state(np.stack([z] * C), np.stack([y] * C))
            eng.set_sampler(sc)
            eng.connectivity(fetch=False)
            eng.sweep(1)
            zz, yy = eng.get_state()
            S_inc = eng.get_connectivity()
            S_new = eng.connectivity()
            out[prec] = (yy, S_inc, S_new)
    y64, S64, _ = out[mb.FP64]
    y32, S32, S32new = out[mb.FP32]
    cand = ((z[:-1] & z[1:]) == 1).sum() * C
    assert (y64 != y32).sum() <= max(2, 2e-3 * cand)
    assert ((y32 <= z[None, :-1]) & (y32 <= z[None, 1:])).all()
    rel_close(S32, S32new, 1e-6, floor=1e-9)                     # two-float rank-1 updates vs from scratch
    same = (y64 == y32).all(axis=2)                               # years whose states coincide: S must agree
    rel_close(S32[same], S64[same], 1e-5, floor=1e-9)


@pytest.mark.parametrize("layout,alpha", [("clusters", 1 / 400), ("blob", 1 / 400), ("uniform", 1 / 20000), ("uniform", 1 / 100),
                                          ("duplicates", 1 / 400), ("line", 1 / 400)])
def test_culled_scan_on_adversarial_landscapes(layout, alpha):
    """The spatial culling (group bounds, reach tests) must stay exact whatever the geometry: tight clusters far
    apart, everything within one dispersal length, no culling at all (tiny alpha), almost everything culled (large
    alpha), many patches on the same spot, patches on a line.  FP32 culled scan vs FP64 generic scan from the same
    state: same decisions up to FP32 ties, incremental S == recomputed S."""
    rng = np.random.default_rng(sum(map(ord, layout)) * 1000 + round(1 / alpha) % 997)
    n, T, C = 4500, 4, 2
    spec, z, y = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.5, miss=0.03)
    if layout == "clusters":
        cen = rng.uniform(0, 30000, (20, 2)); which = rng.integers(0, 20, n)
        spec["px"] = cen[which, 0] + rng.normal(0, 300, n); spec["py"] = cen[which, 1] + rng.normal(0, 300, n)
    elif layout == "blob":
        spec["px"] = rng.uniform(0, 500, n); spec["py"] = rng.uniform(0, 500, n)
    elif layout == "duplicates":
        spots = rng.uniform(0, 20000, (300, 2)); which = rng.integers(0, 300, n)
        spec["px"] = spots[which, 0].copy(); spec["py"] = spots[which, 1].copy()
    elif layout == "line":
        spec["px"] = rng.uniform(0, 400000, n); spec["py"] = np.full(n, 123.0)
    m = O.Model(spec["obs"], geom=O.GEOM_COORDS, px=spec["px"], py=spec["py"], area=spec.get("area"))
    S0 = O.connectivity(m, alpha, 0.5, y[0])
    c = 0.3 / max(np.median(S0[S0 > 0]) if (S0 > 0).any() else 1.0, 1e-300)          # colonisation probabilities around 0.3
    par = pdict(e=0.35, c=float(c), alpha=alpha, b=0.5)
    sc = mb.engine.sampler_config(sample_e=0, sample_c=0, update_z=0, n_adapt=0)
    out = {}
    for prec in (mb.FP64, mb.FP32):
        with make_engine(spec, n_chains=C, precision=prec, seed=23, max_draws=1) as eng:
            eng.set_params([par] * C)
            eng.set_state(np.stack([z] * C), np.stack([y] * C))
            eng.set_sampler(sc)
            eng.connectivity(fetch=False)
            eng.sweep(1)
            _, yy = eng.get_state()
            out[prec] = (yy, eng.get_connectivity(), eng.connectivity())
    y64, S64, _ = out[mb.FP64]
    y32, S32, S32new = out[mb.FP32]
    cand = ((z[:-1] & z[1:]) == 1).sum() * C
    assert (y64 != y32).sum() <= max(2, 2e-3 * cand)
    scale = max(float(np.median(S64[S64 > 0])), 1e-300)
    rel_close(S32, S32new, 1e-6, floor=1e-6 * scale)
    same = (y64 == y32).all(axis=2)
    # FP32 positions: a coordinate is stored to 2^-24 of the landscape's extent, so a weight exp(-alpha d) carries a
    # relative error of up to ~4 alpha extent 2^-24 on top of the FP32 arithmetic (1e-5) -- DESIGN.md section 5
    extent = max(np.ptp(spec["px"]), np.ptp(spec["py"]), 1.0)
    rel_close(S32[same], S64[same], 1e-5 + 4 * alpha * extent * 2.0 ** -24, floor=1e-6 * scale)


def test_fast_sweep_variants_agree(monkeypatch):
    """The launch geometry of the fast sweep (threads per task, cluster size, in-CTA pre-reduction for
    the large-N variants) must not change what a chain does: the Philox thresholds are keyed by cell,
    so different geometries agree except where a log-odds ties with its threshold in FP32."""
    rng = np.random.default_rng(4242)
    n, T, C = 9000, 5, 3
    spec, z, y = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.5, miss=0.03)
    par = pdict(e=0.35, c=0.004, alpha=1 / 400, b=0.5)
    sc = mb.engine.sampler_config(sample_e=0, sample_c=0, update_z=0, n_adapt=0)
    outs = {}
    for tpt, cs in ((512, 4), (1024, 8), (2048, 8), (8192, 8), (512, 1)):
        monkeypatch.setenv("MP_FAST_TPT", str(tpt)); monkeypatch.setenv("MP_FAST_CS", str(cs))
        with make_engine(spec, n_chains=C, precision=mb.FP32, seed=5, max_draws=1) as eng:
            eng.set_params([par] * C)
            eng.set_state(np.stack([z] * C), np.stack([y] * C))
            eng.set_sampler(sc)
            eng.connectivity(fetch=False)
            eng.sweep(1)
            outs[(tpt, cs)] = (eng.get_state()[1], eng.get_connectivity(), eng.connectivity())
    ref_y, ref_S, _ = outs[(512, 4)]
    cand = ((z[:-1] & z[1:]) == 1).sum() * C
    for key, (yy, S_inc, S_new) in outs.items():
        assert (yy != ref_y).sum() <= max(2, 5e-4 * cand), key
        rel_close(S_inc, S_new, 1e-6, floor=1e-9)
        assert ((yy <= z[None, :-1]) & (yy <= z[None, 1:])).all()


def test_large_landscape_sweep_is_consistent():
    """N = 40,000 patches (beyond one CTA's shared memory: cluster of 8 x 256 threads, in-CTA
    pre-reduction): after a sweep the incrementally updated S equals a fresh recomputation and the
    state stays feasible."""
    rng = np.random.default_rng(99)
    n, T = 40000, 3
    spec, z, y = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.45, miss=0.02, areas=False)
    par = pdict(e=0.3, c=0.004, alpha=1 / 400)
    with make_engine(spec, n_chains=1, precision=mb.FP32, seed=8, max_draws=2) as eng:
        eng.set_params([par])
        eng.set_state(z[None], y[None])
        eng.set_sampler(mb.engine.sampler_config(n_adapt=0))
        eng.connectivity(fetch=False)
        eng.sweep(2)
        zz, yy = eng.get_state()
        S_inc = eng.get_connectivity(); S_new = eng.connectivity()
        d = eng.get_draws()
    rel_close(S_inc, S_new, 1e-6, floor=1e-9)
    assert ((yy[0] <= zz[0][:-1]) & (yy[0] <= zz[0][1:])).all()
    assert np.isfinite(d[:, :, 5]).all() and (yy != y[None]).sum() > 1000


def test_cluster_is_widened_until_the_targets_fit_shared_memory():
    """N = 20,000 with 120 (chain, year) tasks: 1,024 threads per task, and the cost model alone would pick one CTA per task
    (120 tasks on 148 SMs) whose 20 slots x 1,024 threads x 16 B do not fit an SM -- the launcher must widen the cluster
    instead of refusing the size.  The sweep runs, S stays consistent with a recomputation, the state stays feasible."""
    rng = np.random.default_rng(2020)
    n, T, C = 20000, 16, 8
    spec, z, y = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.45, miss=0.02, areas=False)
    par = pdict(e=0.3, c=0.004, alpha=1 / 400)
    with make_engine(spec, n_chains=C, precision=mb.FP32, seed=8, max_draws=1) as eng:
        eng.set_params([par] * C)
        eng.set_state(np.stack([z] * C), np.stack([y] * C))
        eng.set_sampler(mb.engine.sampler_config(n_adapt=0))
        eng.connectivity(fetch=False)
        eng.sweep(1)
        geo = eng.scan_geometry()
        zz, yy = eng.get_state()
        S_inc = eng.get_connectivity(); S_new = eng.connectivity()
    assert geo["threads_per_task"] == 1024 and geo["cluster"] >= 2 and geo["culled"], geo
    rel_close(S_inc, S_new, 1e-6, floor=1e-9)
    assert ((yy <= zz[:, :-1]) & (yy <= zz[:, 1:])).all() and (yy != np.stack([y] * C)).sum() > 1000


@pytest.mark.parametrize("n", [3000, 5200])          # 5200: the culled scan and the culled k_conn run sharded
def test_sharded_chain_equals_single_engine(n):
    """BASELINE config 5 path: one chain sharded over W ranks (connectivity by target patches, y scan by
    years, replicated decisions) must reproduce the single-engine run bit for bit.  The W ranks are
    emulated in one process on one GPU, with the NCCL all-reduce replaced by an explicit sum."""
    import torch
    from midaspom_b200 import distributed as D
    rng = np.random.default_rng(31)
    T, C, W = 7, 2, 3
    spec, z, y = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.5, miss=0.05)
    par = pdict(e=0.4, c=0.01, alpha=1 / 400, b=0.5)
    kw = dict(sample_alpha=1, sample_b=1, c_max=0.2, alpha_min=1e-4, alpha_max=1e-1, n_adapt=4)
    nsw = 6

    def fresh():
        eng = make_engine(spec, n_chains=C, precision=mb.FP32, seed=17, max_draws=nsw)
        eng.set_params([par] * C)
        eng.init_chains(mb.engine.sampler_config(**kw), disperse=False)
        return eng

    ref = fresh()
    ref.sweep(nsw)
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


def test_fp64_sweeps_follow_the_cpu_twin_beyond_shared_memory():
    """14,500 patches: the FP64 parity scan keeps its per-task state in global scratch (one CTA's shared memory ends at about
    13,600) and still follows the CPU twin draw by draw."""
    rng = np.random.default_rng(8)
    n, T, C = 14500, 3, 1
    spec, z, _ = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.5, miss=0.05)
    m = make_model(spec)
    kw = dict(sample_alpha=1, sample_b=1, alpha_min=1e-4, alpha_max=5e-2, c_max=1.0, n_adapt=2)
    par0 = pdict(e=0.4, c=0.01, alpha=1 / 400, b=0.5)
    ch = O.Chains(m, O.sampler_cfg(**kw), C, seed=99, par0=oparams(par0), disperse=False)
    want = ch.run(2)
    with make_engine(spec, n_chains=C, seed=99, max_draws=2) as eng:
        eng.set_params([par0] * C)
        eng.init_chains(mb.engine.sampler_config(**kw), disperse=False)
        eng.sweep(2)
        got = eng.get_draws()
        zg, yg = eng.get_state()
    assert (zg == ch.z).all() and (yg == ch.y).all()
    rel_close(got[:, :, :5], want[:, :, :5], 1e-9)
    rel_close(got[:, :, 5], want[:, :, 5], 1e-9, floor=1.0)


@pytest.mark.parametrize("variant", ["dieoff", "loss"])
def test_fp64_variant_sampling_follows_the_cpu_twin(variant):
    """Pre-event years with the die-off scaling K (dieoff.c:56-57,78) or the external source K_L, d_L
    (loss.c:93-101,365), the variant parameters SAMPLED: FP64 engine == CPU twin, draw by draw."""
    rng = np.random.default_rng(123)
    n, T, C = 80, 8, 3
    spec, z, _ = random_landscape(rng, n, T, O.GEOM_LINEAR, occ=0.5, miss=0.06, areas=False)
    spec["era"] = (np.arange(T - 1) < 4).astype(np.uint8)
    m = make_model(spec)
    if variant == "dieoff":
        kw = dict(sample_K=1, K_min=0.1, K_max=100.0)
        par0 = pdict(e=0.4, c=0.2, alpha=1 / 400, K=2.0)
    else:
        kw = dict(sample_Ksrc=1, sample_dsrc=1, Ksrc_min=0.1, Ksrc_max=100.0, dsrc_min=5.0, dsrc_max=400.0)
        par0 = pdict(e=0.4, c=0.2, alpha=1 / 400, K=1.0, Ksrc=2.0, dsrc=60.0)
    kw.update(c_max=2.0, n_adapt=6, n_v_steps=2)
    nsw, seed = 10, 77
    ch = O.Chains(m, O.sampler_cfg(**kw), C, seed=seed, par0=oparams(par0), disperse=False)
    want = ch.run(nsw)
    with make_engine(spec, n_chains=C, seed=seed, max_draws=nsw) as eng:
        eng.set_params([par0] * C)
        eng.init_chains(mb.engine.sampler_config(**kw), disperse=False)
        eng.sweep(nsw)
        got = eng.get_draws()
        zg, yg = eng.get_state()
    assert (zg == ch.z).all() and (yg == ch.y).all()
    rel_close(got[:, :, [0, 1, 8, 9, 10]], want[:, :, [0, 1, 8, 9, 10]], 1e-9)
    rel_close(got[:, :, 5], want[:, :, 5], 1e-9, floor=1.0)
    moved = 8 if variant == "dieoff" else 9
    assert np.unique(got[:, :, moved]).size > 3                     # the variant parameter actually moves


def test_scan_order_equals_the_oracle_rule():
    """mp_get_scan_order == spom_scan_order: both sides visit the patches of a year in the same order."""
    rng = np.random.default_rng(77)
    for geom, n in ((O.GEOM_COORDS, 1), (O.GEOM_COORDS, 37), (O.GEOM_COORDS, 5003), (O.GEOM_LINEAR, 300), (O.GEOM_DENSE, 64)):
        spec, z, y = random_landscape(rng, n, 3, geom)
        if geom == O.GEOM_COORDS and n > 40:
            spec["px"][7] = spec["px"][21]; spec["py"][7] = spec["py"][21]      # a tie
        m = O.Model(spec["obs"], geom=geom, spacing=spec.get("spacing", 100.0), px=spec.get("px"), py=spec.get("py"),
                    dist=spec.get("dist"), area=spec.get("area"))
        with make_engine(spec, n_chains=1, precision=mb.FP32) as eng:
            assert (eng.scan_order() == O.scan_order(m)).all()


def test_chain_offset_selects_the_stream():
    """A chain's random stream depends on its GLOBAL id only: chains [2,3] run alone reproduce
    chains 2,3 of a 4-chain engine (the property MIDASPOM_MPI's row split relies on, :361-372)."""
    rng = np.random.default_rng(9)
    spec, z, _ = random_landscape(rng, 64, 6, O.GEOM_COORDS, miss=0.1, areas=False)
    sc = mb.engine.sampler_config(n_adapt=5)
    outs = []
    for C, off in ((4, 0), (2, 2)):
        with make_engine(spec, n_chains=C, seed=7, max_draws=8, chain_offset=off, precision=mb.FP32) as eng:
            eng.set_params([pdict(e=0.4, c=0.03)] * C)
            eng.init_chains(sc, disperse=True)
            eng.sweep(8)
            outs.append((eng.get_draws(), eng.get_state()))
    assert (outs[0][0][:, 2:, :] == outs[1][0]).all()
    assert (outs[0][1][1][2:] == outs[1][1][1]).all()


# ------------------------------------------------------------------ sampler: posterior vs the reference's exact grid
def test_fp32_sampler_posterior_matches_reference_grid(golden, example_obs):
    """BASELINE north_star: posterior agreement is statistical.  (e, c) draws of the FP32 engine on
    the bundled example vs the exact 101x101 posterior written by MIDASPOM.out (run_examples.sh:8)."""
    spec = dict(obs=example_obs, spacing=100.0, prior_occ=0.5)
    C, nsw, burn = 8, 12000, 1000
    with make_engine(spec, n_chains=C, precision=mb.FP32, seed=99, max_draws=nsw) as eng:
        eng.set_params([pdict(alpha=A)] * C)
        eng.init_chains(mb.engine.sampler_config(n_adapt=500, n_c_steps=2), disperse=True)
        eng.sweep(nsw)
        d = eng.get_draws()[burn:]
    grid, pe, pc, w = grid_marginals(golden["post_default"])
    for col, pm in ((0, pe), (1, pc)):
        mean, sd = grid_moments(grid, pm)
        x = d[:, :, col]
        n_eff = sum(ess(x[:, i]) for i in range(C))
        assert abs(x.mean() - mean) < 5 * sd / np.sqrt(n_eff) + 2e-3
        assert abs(x.std() - sd) < 0.06 * sd
        assert rhat(x) < 1.02
        lo, hi = np.quantile(x, [0.025, 0.975])
        cdf = np.cumsum(pm)
        glo, ghi = grid[np.searchsorted(cdf, 0.025)], grid[np.searchsorted(cdf, 0.975)]
        assert abs(lo - glo) < 0.03 and abs(hi - ghi) < 0.03     # 95% CIs overlap to within 3 grid steps
        thin = max(1, int(np.ceil(x.size / n_eff)) * 2)
        dist, nn = ks_distance_thinned(x.T.ravel(), grid, pm, thin)
        assert dist < 1.63 / np.sqrt(nn) + 0.01


def test_fp32_culled_engine_posterior_matches_fp64_engine():
    """Long-run agreement of the throughput path with the parity path above 3,000 patches, where the FP32 engine runs
    the spatially culled scan, the culled k_conn and refreshes S only every 16th sweep: same data, same sampler
    settings, different seeds -- the posterior means of (e, c, alpha, b) must agree within Monte-Carlo error, the
    posterior widths within 25 %, and every chain of both engines must converge (split R-hat)."""
    from midaspom_b200 import synth
    rng = np.random.default_rng(2024)
    n, T, C, nsw, burn = 3300, 6, 4, 700, 200
    side = np.sqrt(n) * 250.0
    px, py, area = rng.uniform(0, side, n), rng.uniform(0, side, n), rng.lognormal(0, 0.5, n)
    W = synth.kernel_matrix(px, py, area, 1 / 400, 0.5)
    z = np.zeros((T, n), dtype=np.uint8); z[0] = rng.random(n) < 0.5
    cc = None
    for t in range(T - 1):
        yv = z[t] & (rng.random(n) > 0.3)
        S = yv.astype(np.float32) @ W
        cc = cc or 0.3 / S.mean()
        z[t + 1] = np.where(yv == 1, 1, rng.random(n) < np.minimum(1.0, cc * S))
    obs = z.astype(np.int8); hide = rng.random(z.shape) < 0.05; hide[0] = False; obs[hide] = -1
    spec = dict(geom=O.GEOM_COORDS, px=px, py=py, area=area, obs=obs)
    par = pdict(e=0.4, c=float(cc), alpha=1 / 400, b=0.5)
    kw = dict(sample_e=1, sample_c=1, sample_alpha=1, sample_b=1, c_max=20 * float(cc), alpha_min=1e-4, alpha_max=1e-1,
              b_min=0.0, b_max=2.0, n_adapt=150)
    draws = {}
    for prec, seed in ((mb.FP64, 5), (mb.FP32, 6)):
        with make_engine(spec, n_chains=C, precision=prec, seed=seed, max_draws=nsw) as eng:
            eng.set_params([par] * C)
            eng.init_chains(mb.engine.sampler_config(**kw), disperse=False)
            eng.sweep(nsw)
            draws[prec] = eng.get_draws()[burn:]
    for col, name in ((0, "e"), (1, "c"), (2, "alpha"), (3, "b")):
        a, b_ = draws[mb.FP64][:, :, col], draws[mb.FP32][:, :, col]
        na, nb = sum(ess(a[:, i]) for i in range(C)), sum(ess(b_[:, i]) for i in range(C))
        se = np.sqrt(a.var() / max(na, 4) + b_.var() / max(nb, 4))
        assert abs(a.mean() - b_.mean()) < 4.5 * se, (name, a.mean(), b_.mean(), se)
        assert 0.75 < a.std() / b_.std() < 1.33, (name, a.std(), b_.std())
        assert rhat(a) < 1.1 and rhat(b_) < 1.1, (name, rhat(a), rhat(b_))


# ------------------------------------------------------------------ forward simulator (a9)
@pytest.mark.parametrize("geom,n,years,nsims", [(O.GEOM_LINEAR, 50, 12, 6), (O.GEOM_COORDS, 50, 12, 6), (O.GEOM_DENSE, 90, 5, 3),
                                                (O.GEOM_COORDS, 700, 4, 3)])     # 700: several target tiles and survivor-list tiles
def test_simulator_follows_the_cpu_twin(geom, n, years, nsims):
    rng = np.random.default_rng(21)
    spec, z, _ = random_landscape(rng, n, 2, geom, areas=True)
    par = pdict(e=0.5, c=0.05, alpha=1 / 500, b=0.4, K=1.5, Ksrc=0.7, dsrc=60.0)
    spec_o = dict(spec); spec_o["era"] = np.ones(years, dtype=np.uint8)
    m = make_model(spec_o)
    with make_engine(spec, precision=mb.FP64) as eng:
        zs, occ = eng.simulate(par, z[0], years, nsims=nsims, seed=5, era_all=True)
    for s in range(nsims):
        want = O.simulate(m, oparams(par), 5, s, z[0], years)
        assert (zs[s] == want).all()
        assert (occ[s] == want.sum(axis=1)).all()


def test_simulator_one_step_distribution_matches_simpij(golden):
    """Occupancy frequencies after one year vs 10,000 draws of the reference's simpij (libc rand)."""
    z0 = golden["simpij_z0"].astype(np.uint8)
    e, c, K, Ks, a, d, ds = golden["simpij_pars"]
    spec = dict(obs=np.zeros((2, len(z0)), np.int8), spacing=d)
    nsim = 40000
    with make_engine(spec, precision=mb.FP32) as eng:
        zs, _ = eng.simulate(pdict(e=e, c=c, alpha=a, K=K, Ksrc=Ks, dsrc=ds), z0, 1, nsims=nsim, seed=3, era_all=True)
    freq = zs[:, 1, :].mean(axis=0)
    ref = golden["simpij_freq"]
    se = np.sqrt(ref * (1 - ref) * (1 / nsim + 1 / int(golden["simpij_nsim"])))
    assert (np.abs(freq - ref) < 4.5 * se + 1e-9).all()


# ------------------------------------------------------------------ edge cases
def test_edge_shapes():
    """Single patch, two years, all-missing rows, N not a multiple of any tile."""
    for n, T in ((1, 2), (2, 2), (33, 3), (129, 2)):
        rng = np.random.default_rng(n)
        spec, z, y = random_landscape(rng, n, T, O.GEOM_LINEAR, areas=False)
        spec["obs"][1:] = -1                                   # later years entirely unobserved
        m = make_model(spec)
        par = pdict(e=0.4, c=0.3, alpha=A)
        with make_engine(spec, max_draws=4) as eng:
            ll, _ = eng.loglik_host([par], z[None], y[None])
            rel_close(ll[0], O.loglik(m, oparams(par), z, y)[0], 1e-10, floor=1.0)
            eng.set_params([par])
            eng.init_chains(mb.engine.sampler_config(n_adapt=2), disperse=False)
            eng.sweep(4)
            zg, yg = eng.get_state()
            assert ((yg[0] <= zg[0][:-1]) & (yg[0] <= zg[0][1:])).all()
            assert (zg[0][0] == (spec["obs"][0] == 1)).all() or (spec["obs"][0] == -1).any()


# ------------------------------------------------------------------ full-size properties (BASELINE cfg3 shape)
def test_full_size_properties_cfg3(monkeypatch):
    """N=10,000 x T=20 (cfg3 shape), FP32 engine: connectivity is linear in y, the rank-1 log-odds
    equals the difference of two full evaluations, and after a sweep the incrementally updated S
    equals a from-scratch recomputation while the state stays feasible."""
    monkeypatch.setenv("MP_CONN_GEMM", "0")      # properties of k_conn (FP64 accumulation); the tensor-core path: test_gpu_gemm.py
    rng = np.random.default_rng(12345)
    n, T, C = 10000, 20, 2
    spec, z, y = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.4, miss=0.05, areas=True)
    par = pdict(e=0.3, c=0.004, alpha=1 / 400, b=0.5)
    with make_engine(spec, n_chains=C, precision=mb.FP32, max_draws=2, seed=3) as eng:
        eng.set_params([par] * C)
        ya = y * (rng.random(y.shape) < 0.5)
        yb = y - ya
        eng.set_state(np.stack([z, z]), np.stack([ya, yb]))
        Sab = eng.connectivity()
        eng.set_state(np.stack([z, z]), np.stack([y, np.zeros_like(y)]))
        Sy = eng.connectivity()
        # linearity: identical weights, FP32 partial sums per 32 sources (k_conn's FP32 contraction: <= ~1.5e-7 each) joined in FP64
        rel_close(Sab[0] + Sab[1], Sy[0], 1e-6, floor=1e-9)
        assert (Sy[1] == 0).all()
        t, k = 7, int(np.flatnonzero(z[7] & z[8])[5])
        delta32 = eng.flip_delta(0, t, k)
        with make_engine(spec, n_chains=2, precision=mb.FP64) as e64:
            y2 = y.copy(); y2[t, k] ^= 1
            e64.set_params([par] * 2)
            e64.set_state(np.stack([z, z]), np.stack([y, y2]))
            ll, _ = e64.loglik()
            delta64 = e64.flip_delta(0, t, k)
        assert abs((ll[1] - ll[0]) - delta64) < 1e-6           # rank-1 form == difference of full sums
        assert abs(delta32 - delta64) < 1e-3 * max(1.0, abs(delta64))
        # a sweep keeps everything consistent
        eng.set_state(np.stack([z, z]), np.stack([y, y]))
        eng.init_chains(mb.engine.sampler_config(sample_alpha=1, sample_b=1, alpha_min=1e-4, alpha_max=1e-1,
                                                 c_max=0.1, n_adapt=2), disperse=False)
        eng.sweep(2)
        zg, yg = eng.get_state()
        S_inc = eng.get_connectivity()
        S_new = eng.connectivity()
        # the FP32 engine's from-scratch S skips source groups whose weights are below 2^-30 of the target group's S
        # (k_conn, culled variant) while the rank-1 updates reach to 2^-36 of S: the two agree to ~1e-7, any lost update
        # would show at >= 1e-4
        rel_close(S_inc, S_new, 1e-6, floor=1e-9)
        obs = spec["obs"]
        for c in range(C):
            assert ((yg[c] <= zg[c][:-1]) & (yg[c] <= zg[c][1:])).all()
            assert (zg[c][obs == 1] == 1).all() and (zg[c][obs == 0] == 0).all()
        d = eng.get_draws()
        assert np.isfinite(d[:, :, 5]).all()


@pytest.mark.parametrize("geom,n", [(O.GEOM_COORDS, 700), (O.GEOM_LINEAR, 600), (O.GEOM_COORDS, 4100)])
def test_evaluation_calls_between_sweeps_leave_the_sampler_exact(geom, n):
    """mp_loglik / mp_connectivity of an FP32 engine take the FP32 year contraction of k_conn (partial sums per 32 sources,
    ~1e-7 of S) and leave that S resident; the sweep that follows must recompute the resident S with the FP64 contraction
    before the y scan applies rank-1 updates to it -- also when sweeps are replayed as CUDA graphs and the call falls between
    two refreshes.  Checked where it shows: on the linear landscape isolated targets lose orders of magnitude of S when their
    neighbours are removed, and an inexact start would leave a relative error of order one there."""
    rng = np.random.default_rng(900 + n)
    T, C = 6, 3
    spec, z, y = random_landscape(rng, n, T, geom, occ=0.5, miss=0.05)
    par = pdict(e=0.35, c=0.02 if geom != O.GEOM_LINEAR else 0.2, alpha=1 / 400, b=0.5)
    sc = mb.engine.sampler_config(sample_e=1, sample_c=1, sample_alpha=1, sample_b=1, alpha_min=1e-4, alpha_max=1e-1, c_max=0.5, n_adapt=2)
    with make_engine(spec, n_chains=C, precision=mb.FP32, seed=5, max_draws=64) as eng:
        eng.set_params([par] * C)
        eng.set_state(np.stack([z] * C), np.stack([y] * C))
        eng.init_chains(sc, disperse=False)
        eng.sweep(20)                                             # both kinds of sweep have been captured by now
        ll_a, _ = eng.loglik()                                    # FP32 contraction; the resident S is now the evaluation's
        eng.sweep(3)                                              # sweeps 20..22: none of them is a scheduled refresh
        zz, yy = eng.get_state()
        S_inc = eng.get_connectivity()
        ll_b, _ = eng.loglik()
        S_new = eng.get_connectivity()
        rel_close(S_inc, S_new, 1e-6, floor=1e-9)                 # rank-1 bookkeeping on top of an exact start
        assert np.isfinite(ll_a).all() and np.isfinite(ll_b).all()
        assert ((yy <= zz[:, :-1]) & (yy <= zz[:, 1:])).all()
        assert eng.num_draws() == 23
        d = eng.get_draws()
        assert np.isfinite(d[:, :, 5]).all()
