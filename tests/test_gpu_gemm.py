"""The tensor-core connectivity path (k_conn_gemm: tcgen05.mma + TMEM + TMA, midaspom_b200/csrc/mp_conn_gemm.cu) that
mp_connectivity / mp_loglik take when every chain holds the same (alpha, b): S and the log-likelihood against the CPU
oracle at the FP32 tolerance of north_star (1e-5), on ragged sizes and at the cfg3 shape, and against k_conn."""
import numpy as np
import pytest

import midaspom_b200 as mb
import oracle_lib as O
from midaspom_b200 import synth
from gpu_util import make_engine, make_model, pdict, oparams, random_landscape

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_err(got, want, floor):
    got, want = np.asarray(got, float), np.asarray(want, float)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor)))


def oracle_S(m, par, y):
    return np.array([O.connectivity(m, par["alpha"], par["b"], y[t]) for t in range(y.shape[0])])


@pytest.mark.parametrize("geom,n,T,C", [(O.GEOM_COORDS, 1500, 5, 2), (O.GEOM_COORDS, 1061, 9, 5), (O.GEOM_LINEAR, 1300, 4, 3),
                                        (O.GEOM_COORDS, 2048, 12, 9), (O.GEOM_COORDS, 1100, 3, 1)])
def test_gemm_connectivity_vs_oracle_on_ragged_sizes(geom, n, T, C):
    """Sizes that are not multiples of the 128 x 64 tiles; 4 ... 99 columns (every column-block width of the kernel)."""
    rng = np.random.default_rng(n + T)
    spec, z, y = random_landscape(rng, n, T, geom)
    m = make_model(spec)
    par = pdict(alpha=1 / 350, b=0.6 if geom == O.GEOM_COORDS else 0.3)
    ys = np.stack([np.roll(y, c, axis=1) for c in range(C)])
    ys[-1, 0] = 0                                                 # an empty year: exactly zero
    with make_engine(spec, n_chains=C, precision=mb.FP32) as eng:
        eng.set_params([par] * C)
        eng.set_state(np.stack([z] * C), ys)
        S = eng.connectivity()
        assert eng.conn_path() == "gemm"
        assert eng.work_counters()["gemm_tiles"] > 0
    for c in range(C):
        want = oracle_S(m, par, ys[c])
        assert rel_err(S[c], want, 1e-6 * want.max()) <= TOL, (c, rel_err(S[c], want, 1e-6 * want.max()))
    assert (S[-1, 0] == 0).all()


def test_gemm_is_taken_only_for_shared_parameters(monkeypatch):
    rng = np.random.default_rng(5)
    spec, z, y = random_landscape(rng, 1200, 4, O.GEOM_COORDS)
    with make_engine(spec, n_chains=2, precision=mb.FP32) as eng:
        eng.set_state(np.stack([z] * 2), np.stack([y] * 2))
        eng.set_params([pdict(alpha=1 / 400, b=0.5), pdict(alpha=1 / 300, b=0.5)])
        S_a = eng.connectivity()
        assert eng.conn_path() == "k_conn"                         # different alpha: per-chain kernel matrices
        eng.set_params([pdict(alpha=1 / 400, b=0.5)] * 2)
        S_b = eng.connectivity()
        assert eng.conn_path() == "gemm"
    assert rel_err(S_b[0], S_a[0], 1e-6 * S_a[0].max()) <= TOL     # chain 0 has the same parameters in both calls
    with make_engine(spec, n_chains=2, precision=mb.FP64) as eng:  # the FP64 parity engine never takes it
        eng.set_state(np.stack([z] * 2), np.stack([y] * 2))
        eng.set_params([pdict(alpha=1 / 400, b=0.5)] * 2)
        eng.connectivity(fetch=False)
        assert eng.conn_path() == "k_conn"
    monkeypatch.setenv("MP_CONN_GEMM", "0")
    with make_engine(spec, n_chains=2, precision=mb.FP32) as eng:
        eng.set_state(np.stack([z] * 2), np.stack([y] * 2))
        eng.set_params([pdict(alpha=1 / 400, b=0.5)] * 2)
        eng.connectivity(fetch=False)
        assert eng.conn_path() == "k_conn"


def test_gemm_at_cfg3_shape_vs_oracle_and_k_conn(monkeypatch):
    """N=10,000 x T=20, 8 chains sharing the true (alpha, b): 152 columns = one 160-wide column block."""
    wl = synth.make_workload("cfg3")
    t = wl["truth"]
    C = 8
    par = pdict(e=t["e"], c=0.5 * t["c"], alpha=t["alpha"], b=t["b"])
    rng = np.random.default_rng(3)
    z = wl["z_true"].astype(np.uint8)
    ys = np.stack([(z[:-1] & z[1:] & (rng.random((wl["T"] - 1, wl["n"])) < 0.7)).astype(np.uint8) for _ in range(C)])
    spec = dict(geom=O.GEOM_COORDS, px=wl["px"], py=wl["py"], area=wl["area"], obs=wl["obs"])
    m = make_model(spec)
    with make_engine(spec, n_chains=C, precision=mb.FP32) as eng:
        ll, parts = eng.loglik_host([par] * C, np.stack([z] * C), ys)
        assert eng.conn_path() == "gemm"
        S = eng.get_connectivity()
    monkeypatch.setenv("MP_CONN_GEMM", "0")
    with make_engine(spec, n_chains=C, precision=mb.FP32) as eng:
        ll_k, parts_k = eng.loglik_host([par] * C, np.stack([z] * C), ys)
        assert eng.conn_path() == "k_conn"
        S_k = eng.get_connectivity()
    assert rel_err(S, S_k, 1e-6 * S_k.max()) <= TOL
    assert np.max(np.abs(ll - ll_k) / np.abs(ll_k)) <= TOL
    for c in (0, C - 1):
        want, wparts, wS = O.loglik(m, oparams(par), z, ys[c], want_S=True)
        assert rel_err(S[c], wS, 1e-6 * wS.max()) <= TOL, rel_err(S[c], wS, 1e-6 * wS.max())
        assert abs(ll[c] - want) <= TOL * abs(want)


def test_gemm_cfg5_sampled_targets_and_linearity():
    """N=100,000 x T=30, one chain: 29 columns on the tensor cores against the oracle on 256 sampled targets; and the
    contraction is linear in y to the FP32 accumulation error."""
    wl = synth.make_workload("cfg5")
    n, T = wl["n"], wl["T"]
    t = wl["truth"]
    par = pdict(e=t["e"], c=0.5 * t["c"], alpha=t["alpha"], b=t["b"])
    rng = np.random.default_rng(3)
    z = wl["z_true"].astype(np.uint8)
    y = (z[:-1] & z[1:] & (rng.random((T - 1, n)) < 0.7)).astype(np.uint8)
    ya = (y * (rng.random(y.shape) < 0.5)).astype(np.uint8)
    spec = dict(geom=O.GEOM_COORDS, px=wl["px"], py=wl["py"], area=wl["area"], obs=wl["obs"])
    m = make_model(spec)
    with make_engine(spec, n_chains=1, precision=mb.FP32) as eng:
        eng.set_params([par])
        eng.set_state(z[None], y[None])
        S = eng.connectivity()[0]
        assert eng.conn_path() == "gemm"
        eng.set_state(z[None], ya[None]); Sa = eng.connectivity()[0]
        eng.set_state(z[None], (y - ya)[None]); Sb = eng.connectivity()[0]
    targets = np.sort(rng.choice(n, 256, replace=False)).astype(np.int32)
    for year in (0, T // 2, T - 2):
        want = O.connectivity_targets(m, par["alpha"], par["b"], y[year], targets)
        assert rel_err(S[year, targets], want, 1e-6 * want.max()) <= TOL
    assert rel_err(Sa + Sb, S, 1e-6 * S.max()) <= TOL
