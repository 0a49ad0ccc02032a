"""GPU parity at the shapes BASELINE.json names: connectivity S and the complete-data log-likelihood (with its
four parts) of BOTH engines (FP64 1e-9, FP32 1e-5 -- north_star's tolerances) against the CPU oracle on the very
workloads bench.py times: cfg2 (N=1,000 x T=10, imperfect detection), cfg3 (N=10,000 x T=20, coordinates + areas,
b=0.5), cfg4 / cfg4l (die-off and patch-loss eras at N=10,000 x T=20) in full, and cfg5 (N=100,000 x T=30) on 256
sampled target patches (the oracle needs O(N) per target).

cfg2-cfg5 use extensions the reference has no code for (planar coordinates, areas with exponent b, detection):
the oracle is the checker here, and it is pinned to the reference only at the degenerate point
(tests/test_oracle_vs_reference.py) -- "parity unpinned" for these shapes, stated in DESIGN.md."""
import numpy as np
import pytest

import midaspom_b200 as mb
import oracle_lib as O
from midaspom_b200 import synth
from gpu_util import make_engine, make_model, pdict, oparams

pytestmark = pytest.mark.gpu

RTOL = {mb.FP64: 1e-9, mb.FP32: 1e-5}


def rel_err(got, want, floor):
    got, want = np.asarray(got, float), np.asarray(want, float)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor)))


def workload_state(wl, seed=3):
    """A feasible complete-data state of the workload: the true occupancies and an intermediate state below them."""
    rng = np.random.default_rng(seed)
    z = wl["z_true"].astype(np.uint8)
    y = (z[:-1] & z[1:] & (rng.random((wl["T"] - 1, wl["n"])) < 0.7)).astype(np.uint8)
    return z, y


def workload_spec(wl):
    spec = dict(geom=O.GEOM_COORDS, px=wl["px"], py=wl["py"], area=wl["area"], obs=wl["obs"], detect=wl["detect"])
    if "era" in wl:
        spec["era"] = wl["era"]
    if "src_unit" in wl:
        spec["src_unit"] = wl["src_unit"]
    return spec


_cache = {}


def oracle_values(name):
    """(workload, spec, par, z, y, oracle loglik, parts, S) -- one oracle evaluation per workload for both precisions."""
    if name not in _cache:
        wl = synth.make_workload(name)
        t = wl["truth"]
        # half the true c keeps every colonisation probability below the clamp, so that all terms are finite
        par = pdict(e=t["e"], c=0.5 * t["c"], alpha=t["alpha"], b=t["b"], p=t["p"], K=t.get("K", 1.0), Ksrc=t.get("Ksrc", 0.0),
                    dsrc=t.get("dsrc", 0.0))
        z, y = workload_state(wl)
        spec = workload_spec(wl)
        want, parts, S = O.loglik(make_model(spec), oparams(par), z, y, want_S=True)
        _cache[name] = (wl, spec, par, z, y, want, parts, S)
    return _cache[name]


@pytest.mark.parametrize("precision", [mb.FP64, mb.FP32], ids=["fp64", "fp32"])
@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "cfg4l"])
def test_connectivity_and_loglik_at_baseline_shapes(name, precision, monkeypatch):
    monkeypatch.setenv("MP_CONN_GEMM", "0")      # one chain trivially "shares" (alpha, b): keep k_conn here (the tensor-core path: test_gpu_gemm.py)
    wl, spec, par, z, y, want, wparts, wS = oracle_values(name)
    assert np.isfinite(want)
    tol = RTOL[precision]
    with make_engine(spec, n_chains=1, precision=precision) as eng:
        ll, parts = eng.loglik_host([par], z[None], y[None])
        S1 = eng.get_connectivity()[0]
        # second evaluation: the FP32 engine now culls k_conn against the resident S (the first one has no bound to cull with)
        S2 = eng.connectivity()[0]
        ll2, parts2 = eng.loglik()
        work = eng.work_counters()
    floor = 1e-6 * float(wS.max())
    assert rel_err(S1, wS, floor) <= tol, (name, "S unculled", rel_err(S1, wS, floor))
    assert rel_err(S2, wS, floor) <= tol, (name, "S culled", rel_err(S2, wS, floor))
    for got_ll, got_parts in ((ll, parts), (ll2, parts2)):
        assert rel_err(got_parts[0], wparts, 1.0) <= tol, (name, got_parts[0], wparts)
        assert abs(got_ll[0] - want) <= tol * abs(want), (name, got_ll[0], want)
    if precision == mb.FP32 and wl["n"] >= 10000:
        assert work["conn_exec"] < work["conn_total"]           # the culled kernel really skipped source tiles


@pytest.mark.parametrize("precision", [mb.FP64, mb.FP32], ids=["fp64", "fp32"])
def test_connectivity_at_cfg5_on_sampled_targets(precision, monkeypatch):
    """N=100,000 x T=30: S of 256 sampled target patches in three sampled years against the oracle."""
    monkeypatch.setenv("MP_CONN_GEMM", "0")
    wl = synth.make_workload("cfg5")
    n, T = wl["n"], wl["T"]
    t = wl["truth"]
    par = pdict(e=t["e"], c=0.5 * t["c"], alpha=t["alpha"], b=t["b"])
    z, y = workload_state(wl)
    spec = workload_spec(wl)
    m = make_model(spec)
    rng = np.random.default_rng(11)
    tol = RTOL[precision]
    with make_engine(spec, n_chains=1, precision=precision) as eng:
        eng.set_params([par])
        eng.set_state(z[None], y[None])
        if precision == mb.FP64:
            # the FP64 parity engine never culls: evaluate a target range only (the patch-sharding knob), sample inside it
            # (a range of scan-order slots on landscapes with positions: mp_set_shard)
            lo, hi = 25600, 25600 + 2560
            eng.set_shard(lo, hi, 0, 1)
            S = eng.connectivity()[0]
            targets = np.sort(rng.choice(eng.scan_order()[lo:hi], 256, replace=False)).astype(np.int32)
            checks = [("fp64", S)]
        else:
            S1 = eng.connectivity()[0]
            S2 = eng.connectivity()[0]                           # culled against the resident S
            targets = np.sort(rng.choice(n, 256, replace=False)).astype(np.int32)
            checks = [("fp32 unculled", S1), ("fp32 culled", S2)]
    for year in (0, T // 2, T - 2):
        want = O.connectivity_targets(m, par["alpha"], par["b"], y[year], targets)
        floor = 1e-6 * float(want.max())
        for label, S in checks:
            err = rel_err(S[year, targets], want, floor)
            assert err <= tol, (label, year, err)
