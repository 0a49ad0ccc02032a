"""Builds the same problem twice: as an oracle Model (CPU checker) and as a midaspom_b200.Engine."""
import numpy as np

import midaspom_b200 as mb
import oracle_lib as O


def make_engine(spec, n_chains=1, precision=mb.FP64, seed=1, max_draws=0, chain_offset=0):
    obs = np.asarray(spec["obs"], dtype=np.int8)
    T, n = obs.shape
    eng = mb.Engine(n, T, n_chains, precision=precision, seed=seed, prior_occ=spec.get("prior_occ", 0.5),
                    detect=spec.get("detect", 0), max_draws=max_draws, chain_offset=chain_offset)
    geom = spec.get("geom", O.GEOM_LINEAR)
    if geom == O.GEOM_LINEAR:
        eng.set_landscape_linear(spec.get("spacing", 100.0), spec.get("area"))
    elif geom == O.GEOM_COORDS:
        eng.set_landscape_coords(spec["px"], spec["py"], spec.get("area"))
    else:
        eng.set_landscape_dense(spec["dist"], spec.get("area"))
    eng.set_source_units(spec.get("src_unit"))
    eng.set_observations(obs)
    eng.set_era(spec.get("era"))
    return eng


def make_model(spec):
    return O.Model(spec["obs"], geom=spec.get("geom", O.GEOM_LINEAR), spacing=spec.get("spacing", 100.0),
                   prior_occ=spec.get("prior_occ", 0.5), detect=spec.get("detect", 0), px=spec.get("px"),
                   py=spec.get("py"), dist=spec.get("dist"), area=spec.get("area"), src_unit=spec.get("src_unit"),
                   era=spec.get("era"))


def pdict(**kw):
    d = dict(mb.engine.PARAM_DEFAULTS)
    d.update(kw)
    return d


def oparams(d):
    return O.params(**{k: d[k] for k in ("e", "c", "alpha", "b", "p", "K", "Ksrc", "dsrc")})


def random_landscape(rng, n, T, geom, occ=0.45, miss=0.0, areas=True, side=None):
    side = side or np.sqrt(n) * 250.0
    spec = dict(geom=geom)
    px, py = rng.uniform(0, side, n), rng.uniform(0, side, n)
    if geom == O.GEOM_COORDS:
        spec.update(px=px, py=py)
    elif geom == O.GEOM_DENSE:
        d = np.sqrt((px[:, None] - px[None, :]) ** 2 + (py[:, None] - py[None, :]) ** 2)
        spec.update(dist=d.astype(np.float32).astype(np.float64))   # float-representable: same input in both precisions
    else:
        spec.update(spacing=100.0)
    if areas:
        spec["area"] = rng.lognormal(0, 0.5, n)
    z = (rng.random((T, n)) < occ).astype(np.uint8)
    y = (z[:-1] & z[1:] & (rng.random((T - 1, n)) < 0.6)).astype(np.uint8)
    obs = z.astype(np.int8)
    if miss > 0:
        obs[rng.random((T, n)) < miss] = -1
    spec["obs"] = obs
    return spec, z, y
