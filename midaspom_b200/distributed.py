"""Chain sharding over the GPUs of one box and the posterior gather.

Replaces MIDASPOM_MPI's static block split of grid rows with one MPI_Send/MPI_Recv of result rows
(main_MIDASPOM_MPI.c:192,361-372,483-505): chains are the independent unit, rank r owns the global
chains [r*cpg, (r+1)*cpg) (its Philox streams are keyed by the global id), nothing is exchanged
while sampling, and one all_gather of the draws (NCCL on GPUs, gloo in CPU tests) gives every rank
all chains for R-hat / ESS.
"""
from __future__ import annotations

import numpy as np


def chain_block(rank: int, world: int, chains_per_rank: int):
    """(global id of the first local chain, local chain count)."""
    return rank * chains_per_rank, chains_per_rank


def split_rows(total: int, world: int, rank: int):
    """The reference's row split: rank 0 takes avg + total % world, the others avg
    (main_MIDASPOM_MPI.c:361-372) -> (start, end)."""
    avg = total // world
    if rank == 0:
        return 0, avg + total % world
    start = avg + total % world + (rank - 1) * avg
    return start, start + avg


def gather_draws(draws, group=None):
    """draws: torch tensor (sweeps, local_chains, fields) on this rank's device -> (sweeps, all chains, fields)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return draws
    world = dist.get_world_size(group)
    parts = [torch.empty_like(draws) for _ in range(world)]
    dist.all_gather(parts, draws.contiguous(), group=group)
    return torch.cat(parts, dim=1)


def split_rhat(x: np.ndarray) -> float:
    """Split R-hat of draws x chains."""
    x = np.asarray(x, dtype=float)
    n = x.shape[0] // 2
    if n < 2:
        return float("nan")
    h = np.concatenate([x[:n], x[n:2 * n]], axis=1)
    W = h.var(axis=0, ddof=1).mean()
    B = n * h.mean(axis=0).var(ddof=1)
    return float(np.sqrt(((n - 1) / n * W + B / n) / W)) if W > 0 else 1.0


def ess_geyer(x: np.ndarray) -> float:
    """Effective sample size of one chain (Geyer initial positive sequence)."""
    x = np.asarray(x, dtype=float)
    n = len(x)
    if n < 8 or not np.isfinite(x).all() or x.var() <= 0.0 or (x - x.mean()).var() <= 0.0:
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


def posterior_summary(all_draws: np.ndarray, fields=("e", "c", "alpha", "b", "p")):
    """all_draws: (sweeps, chains, >=len(fields)) -> per-parameter mean / sd / R-hat / total ESS."""
    out = {}
    for i, f in enumerate(fields):
        x = all_draws[:, :, i]
        if x.std() == 0:
            continue
        out[f] = dict(mean=float(x.mean()), sd=float(x.std()), rhat=split_rhat(x),
                      ess=float(sum(ess_geyer(x[:, c]) for c in range(x.shape[1]))))
    return out
