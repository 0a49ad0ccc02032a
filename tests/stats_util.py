"""Small statistics helpers shared by the CPU and GPU posterior tests (numpy only)."""
import numpy as np


def grid_marginals(post, lo=0.0, hi=1.0):
    """Marginals of the reference's s x s posterior table with its trapezoid weights
    (main_MIDASPOM.c:414-424; rows = e, columns = c, as Rscript/plot_posterior.R:22-35 reads it)."""
    s = post.shape[0]
    grid = np.linspace(lo, hi, s)
    w1 = np.ones(s); w1[0] = w1[-1] = 0.5
    w = np.outer(w1, w1) * post
    w = w / w.sum()
    return grid, w.sum(axis=1), w.sum(axis=0), w


def grid_moments(grid, pm):
    mean = (pm * grid).sum()
    sd = np.sqrt((pm * (grid - mean) ** 2).sum())
    return mean, sd


def grid_cdf(grid, pm, x):
    """CDF of the piecewise-constant density implied by the grid masses (cell centred on each node)."""
    h = grid[1] - grid[0]
    edges = np.concatenate([[grid[0]], (grid[:-1] + grid[1:]) / 2, [grid[-1]]])
    cum = np.concatenate([[0.0], np.cumsum(pm)])
    return np.interp(x, edges, cum)


def ess(x):
    """Effective sample size of one chain (Geyer initial positive sequence)."""
    x = np.asarray(x, dtype=float)
    n = len(x)
    if n < 8 or x.std() == 0:
        return float(n)
    xc = x - x.mean()
    f = np.fft.rfft(xc, 2 * n)
    acf = np.fft.irfft(f * np.conj(f))[:n] / (np.arange(n, 0, -1) * xc.var())
    tau = 1.0
    for k in range(1, n - 1, 2):
        pair = acf[k] + acf[k + 1]
        if pair < 0:
            break
        tau += 2 * pair
    return n / max(tau, 1e-12)


def ks_distance_thinned(draws, grid, pm, thin):
    x = np.sort(np.asarray(draws)[::thin])
    n = len(x)
    F = grid_cdf(grid, pm, x)
    d = max(np.abs(F - np.arange(1, n + 1) / n).max(), np.abs(F - np.arange(0, n) / n).max())
    return d, n


def rhat(chains):
    """Split R-hat over chains (draws x chains)."""
    x = np.asarray(chains, dtype=float)
    n = x.shape[0] // 2
    halves = np.concatenate([x[:n], x[n:2 * n]], axis=1)
    m = halves.shape[1]
    W = halves.var(axis=0, ddof=1).mean()
    B = n * halves.mean(axis=0).var(ddof=1)
    return float(np.sqrt(((n - 1) / n * W + B / n) / W)) if W > 0 else 1.0
