"""Host-side logic that needs no GPU: the posterior diagnostics used by bench.py and the drivers, the
reference's row split, the synthetic workload generator."""
import numpy as np

from midaspom_b200 import distributed as D
from midaspom_b200 import synth


def ar1(rng, n, rho):
    x = np.empty(n)
    x[0] = rng.normal()
    e = rng.normal(size=n) * np.sqrt(1 - rho * rho)
    for i in range(1, n):
        x[i] = rho * x[i - 1] + e[i]
    return x


def test_ess_of_an_ar1_chain_matches_theory():
    """ESS of an AR(1) chain is n (1 - rho) / (1 + rho)."""
    rng = np.random.default_rng(3)
    n = 40000
    for rho in (0.0, 0.5, 0.9):
        got = D.ess_geyer(ar1(rng, n, rho))
        want = n * (1 - rho) / (1 + rho)
        assert abs(got - want) < 0.15 * want, (rho, got, want)
    assert D.ess_geyer(np.ones(100)) == 100.0                     # constant chain: defined as n, never NaN
    assert D.ess_geyer(np.arange(5.0)) == 5.0                     # too short to estimate


def test_split_rhat_separates_converged_from_stuck_chains():
    rng = np.random.default_rng(4)
    good = rng.normal(size=(2000, 4))
    assert D.split_rhat(good) < 1.01
    bad = good.copy(); bad[:, 0] += 3.0                           # one chain somewhere else
    assert D.split_rhat(bad) > 1.5
    drift = good + np.linspace(0, 3, 2000)[:, None]               # all chains drifting: caught by the split
    assert D.split_rhat(drift) > 1.2
    assert np.isnan(D.split_rhat(good[:3]))


def test_posterior_summary_fields_and_constant_parameters():
    rng = np.random.default_rng(5)
    d = rng.normal(size=(500, 3, 5)) * np.array([0.1, 0.2, 0.0, 0.3, 0.0]) + np.array([0.3, 0.05, 0.0025, 0.5, 1.0])
    s = D.posterior_summary(d)
    assert set(s) == {"e", "c", "b"}                              # alpha and p never moved: not reported
    assert abs(s["e"]["mean"] - 0.3) < 0.02 and abs(s["b"]["sd"] - 0.3) < 0.03
    assert s["c"]["ess"] > 1000 and s["c"]["rhat"] < 1.02


def test_reference_row_split_and_chain_blocks():
    """main_MIDASPOM_MPI.c:361-372: rank 0 takes the remainder."""
    for total, world in ((101, 4), (101, 1), (7, 8), (64, 8)):
        spans = [D.split_rows(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert spans[0][1] - spans[0][0] == total // world + total % world
    assert [D.chain_block(r, 8, 8) for r in (0, 3, 7)] == [(0, 8), (24, 8), (56, 8)]


def test_synthetic_workloads_are_deterministic_and_feasible():
    a, b = synth.make_workload("tiny"), synth.make_workload("tiny")
    assert (a["obs"] == b["obs"]).all() and (a["px"] == b["px"]).all() and a["truth"] == b["truth"]
    z, obs = a["z_true"], a["obs"]
    assert obs.shape == (a["T"], a["n"]) and set(np.unique(obs)) <= {-1, 0, 1}
    assert (obs[0] != -1).all()                                   # year 0 is never hidden (UB in the reference)
    assert ((obs == z) | (obs == -1)).all()                       # perfect detection: what is seen is the truth
    w = synth.make_workload("cfg2")                               # imperfect detection: misses only, never false presences
    assert w["detect"] == 1 and ((w["obs"] <= w["z_true"]) | (w["obs"] == -1)).all() and (w["obs"][w["z_true"] == 1] == 0).any()
    W = synth.kernel_matrix(a["px"], a["py"], a["area"], 1 / 400, 0.5, dtype=np.float64)
    assert (np.diag(W) == 0).all()
    i, j = 3, 17
    assert np.isclose(W[i, j], np.exp(-np.hypot(a["px"][i] - a["px"][j], a["py"][i] - a["py"][j]) / 400) * a["area"][i] ** 0.5)
