"""Synthetic landscapes and occupancy histories of the shapes BASELINE.json names.

Pure numpy (host side, not timed): uniform patch positions in a square of side sqrt(N)*250 m
(seed 12345), lognormal(0, 0.5) areas (seed 12346), an occupancy history drawn from the model
itself -- the generative step of the reference's simpij (main_MIDASPOM_future.c:64-110):
survive iff u > E, colonise iff u < min(1, c*S) with S computed from the post-extinction state --
then 5 % of the cells of years >= 1 set to -1 and, for imperfect detection, occupied cells reported
as 0 with probability 1-p (Rscript/simuls_traj.R:107-133).
"""
from __future__ import annotations

import numpy as np

WORKLOADS = {
    # name: (N patches, T years, chains per GPU, imperfect detection)
    "cfg2": dict(n=1000, T=10, chains_per_gpu=8, detect=1, desc="synthetic N=1,000 x T=10, 8 chains, imperfect detection"),
    "cfg3": dict(n=10000, T=20, chains_per_gpu=8, chains_total=64, detect=0, desc="synthetic N=10,000 x T=20, 64 independent chains sharded over the GPUs"),
    "cfg4": dict(n=10000, T=20, chains_per_gpu=8, detect=0, era_pre=10, K=3.0,
                 desc="synthetic N=10,000 x T=20, 8 chains per GPU, die-off variant: first 10 transitions pre-event with K_D=3 (E=e/K, C=c K S), K sampled"),
    "cfg4l": dict(n=10000, T=20, chains_per_gpu=8, detect=0, era_pre=10, Ksrc=5.0, dsrc=500.0,
                  desc="synthetic N=10,000 x T=20, 8 chains per GPU, patch-loss variant: first 10 transitions pre-event with an external "
                       "source K_L=5 at unit distance d_L=500 m west of the landscape (C = c (S + K_L exp(-alpha u_k d_L))), K_L and d_L sampled"),
    "tiny": dict(n=256, T=6, chains_per_gpu=4, detect=0, desc="smoke-test size"),
    "cfg5": dict(n=100000, T=30, chains_per_gpu=1, detect=0, desc="synthetic N=100,000 x T=30, one chain"),
    "cfg5t": dict(n=100000, T=5, chains_per_gpu=1, detect=0, desc="synthetic N=100,000 x T=5, one chain (4 year tasks: what one of 8 GPUs scans in cfg5)"),
    "cfg5s": dict(n=40000, T=12, chains_per_gpu=1, detect=0, desc="synthetic N=40,000 x T=12, one chain (reduced cfg5)"),
}
TRUTH = dict(e=0.3, alpha=1.0 / 400.0, b=0.5, p_detect=0.8, target_mean_C=0.3)


def kernel_matrix(px, py, area, alpha, b, dtype=np.float32):
    """W[l, k] = exp(-alpha d_lk) A_l^b, zero diagonal ([source][target], like M in main_MIDASPOM.c:180-188)."""
    n = len(px)
    W = np.empty((n, n), dtype=dtype)
    aw = (area ** b).astype(dtype)
    step = max(1, (1 << 24) // n)
    for i in range(0, n, step):
        d = np.sqrt((px[i:i + step, None] - px[None, :]) ** 2 + (py[i:i + step, None] - py[None, :]) ** 2)
        W[i:i + step] = np.exp(-alpha * d).astype(dtype) * aw[i:i + step, None]
    np.fill_diagonal(W, 0)
    return W


def make_workload_large(name: str, seed: int = 12345, device: int = 0):
    """Same generative model for landscapes whose N x N kernel matrix does not fit the host (cfg5): the occupancy history
    comes from the engine's own forward simulator (mp_simulate: simpij, main_MIDASPOM_future.c:64-110, with Philox draws),
    after one mp_connectivity call that normalises c like make_workload does.  Needs a GPU."""
    from .engine import Engine, FP32
    w = WORKLOADS[name]
    n, T = w["n"], w["T"]
    rng = np.random.default_rng(seed)
    side = np.sqrt(n) * 250.0
    px, py = rng.uniform(0, side, n), rng.uniform(0, side, n)
    area = np.random.default_rng(seed + 1).lognormal(0.0, 0.5, n)
    alpha, b, e = TRUTH["alpha"], TRUTH["b"], TRUTH["e"]
    eng = Engine(n, 2, 1, precision=FP32, device=device)
    eng.set_landscape_coords(px, py, area)
    eng.set_source_units(None)
    eng.set_params([dict(e=e, c=1.0, alpha=alpha, b=b)])
    z0 = (rng.random(n) < 0.5).astype(np.uint8)
    y0 = z0 & (rng.random(n) > e)
    eng.set_state(np.stack([z0, z0])[None], y0[None, None])
    S0 = eng.connectivity()[0, 0]
    c = float(TRUTH["target_mean_C"] / max(S0.mean(), 1e-30))      # mean colonisation probability of the first transition ~0.3
    zs, _ = eng.simulate(dict(e=e, c=c, alpha=alpha, b=b), z0, T - 1, nsims=1, seed=seed)
    eng.close()
    z = np.ascontiguousarray(zs[0], dtype=np.uint8)
    obs = z.astype(np.int8)
    hide = rng.random(z.shape) < 0.05
    hide[0] = False
    obs[hide] = -1
    truth = dict(e=e, c=c, alpha=alpha, b=b, p=1.0)
    return dict(name=name, n=n, T=T, px=px, py=py, area=area, obs=obs, z_true=z, truth=truth, detect=0,
                chains_per_gpu=w["chains_per_gpu"], desc=w["desc"])


def make_workload(name: str, seed: int = 12345, device: int = 0):
    w = WORKLOADS[name]
    n, T = w["n"], w["T"]
    if n > 20000:
        return make_workload_large(name, seed, device)
    rng = np.random.default_rng(seed)
    side = np.sqrt(n) * 250.0
    px, py = rng.uniform(0, side, n), rng.uniform(0, side, n)
    area = np.random.default_rng(seed + 1).lognormal(0.0, 0.5, n)
    alpha, b, e = TRUTH["alpha"], TRUTH["b"], TRUTH["e"]
    W = kernel_matrix(px, py, area, alpha, b)
    z = np.zeros((T, n), dtype=np.uint8)
    z[0] = rng.random(n) < 0.5
    if not z[0].any():
        z[0, 0] = 1
    # c normalised so that the mean colonisation probability of the first transition is ~0.3
    y0 = z[0] & (rng.random(n) > e)
    S0 = y0.astype(np.float32) @ W
    c = float(TRUTH["target_mean_C"] / max(S0.mean(), 1e-30))
    y = y0
    era_pre, Kv = int(w.get("era_pre", 0)), float(w.get("K", 1.0))
    Ksrc, dsrc = float(w.get("Ksrc", 0.0)), float(w.get("dsrc", 0.0))
    # patch-loss variant: the habitat lost at the event was an external source before it (loss.c:86-105,365); here it lies
    # west of the landscape, patch k at u_k distance units from it
    src_unit = (px + 500.0) / 500.0 if Ksrc else None
    for t in range(T - 1):
        Kt = Kv if t < era_pre else 1.0                      # die-off variant: E = e/K, C = c K S before the event (dieoff.c:56-57,78)
        if t > 0 or Kt != 1.0:
            y = z[t] & (rng.random(n) > min(1.0, e / Kt))
        S = y.astype(np.float32) @ W
        if Ksrc and t < era_pre:
            S = S + Ksrc * np.exp(-alpha * src_unit * dsrc)
        C = np.minimum(1.0, c * Kt * S)
        z[t + 1] = np.where(y == 1, 1, rng.random(n) < C)
    obs = z.astype(np.int8)
    if w["detect"]:
        missed = (z == 1) & (rng.random(z.shape) > TRUTH["p_detect"])
        obs[missed] = 0
    hide = rng.random(z.shape) < 0.05
    hide[0] = False
    obs[hide] = -1
    truth = dict(e=e, c=c, alpha=alpha, b=b, p=TRUTH["p_detect"] if w["detect"] else 1.0)
    out = dict(name=name, n=n, T=T, px=px, py=py, area=area, obs=obs, z_true=z, truth=truth, detect=w["detect"],
               chains_per_gpu=w["chains_per_gpu"], desc=w["desc"])
    if "chains_total" in w:
        out["chains_total"] = w["chains_total"]
    if era_pre:
        out["era"] = (np.arange(T - 1) < era_pre).astype(np.uint8)
        truth["K"] = Kv
    if Ksrc:
        out["src_unit"] = src_unit
        truth["Ksrc"], truth["dsrc"] = Ksrc, dsrc
    return out
