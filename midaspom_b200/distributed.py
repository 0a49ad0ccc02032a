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


# ----------------------------------------------------------------------------- one chain over several GPUs
PH_PROPOSE_CONN, PH_DECIDE_Z, PH_SWEEP_Y, PH_FINISH = 0, 1, 2, 3
BUF_Y, BUF_S, BUF_S_PROP = 2, 3, 5


class _DeviceView:
    """__cuda_array_interface__ over a buffer owned by an mp_engine (zero-copy torch view)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def engine_tensor(eng, which, shape, typestr, device):
    import torch
    ptr, _ = eng.device_ptr(which)
    return torch.as_tensor(_DeviceView(ptr, shape, typestr), device=device)


class ShardedChain:
    """Large-N runs (BASELINE config 5): every rank holds a full replica of the chain(s); per sweep
      * the connectivity (N^2 pair terms) is split by TARGET PATCHES: rank r evaluates the columns
        [lo_r, hi_r) of S / S_prop, an all-reduce(sum) over zero-filled buffers assembles them;
      * the Gibbs scan of y is split by YEARS: the (chain, year) tasks are independent given z, rank r scans the
        tasks r, r + W, ...; the updated rows of y (the occupancy state) and S are exchanged the same way;
      * everything else (Metropolis decisions, z update, counts) is replicated -- identical inputs and identical
        Philox counters give identical decisions on every rank, so the result equals a single-GPU run bit for bit.
    `reduce_fn(tensor)` must sum the tensor over ranks in place (torch.distributed.all_reduce by default)."""

    def __init__(self, eng, rank, world, device, reduce_fn=None):
        import torch
        self.eng, self.rank, self.world = eng, rank, world
        C, T, N = eng.C, eng.T, eng.N
        per = -(-N // world)
        per = -(-per // 256) * 256                      # whole k_conn CTAs (128 threads x 2 targets)
        lo, hi = min(N, rank * per), min(N, (rank + 1) * per)
        eng.set_shard(lo, hi, rank, world)
        self.S = engine_tensor(eng, BUF_S, (C * (T - 1), N), "<f8", device)
        self.Sp = engine_tensor(eng, BUF_S_PROP, (C * (T - 1), N), "<f8", device)
        self.y = engine_tensor(eng, BUF_Y, (C * (T - 1), N), "|u1", device)
        own = (torch.arange(C * (T - 1), device=device) % world) == rank
        self.not_own = ~own
        if reduce_fn is None:
            import torch.distributed as dist
            reduce_fn = lambda t: dist.all_reduce(t)
        self.reduce_fn = reduce_fn
        self.torch = torch

    def describe(self):
        return (f"one chain over {self.world} GPUs: connectivity by target patches, y scan by years, "
                "all-reduce of S and y per sweep (NCCL through torch.distributed)")

    def bytes_per_sweep(self):
        C, T, N = self.eng.C, self.eng.T, self.eng.N
        return 2 * C * (T - 1) * N * 8 + C * (T - 1) * N + (C * (T - 1) * N * 8) / 16   # S_prop + S + y every sweep; refresh of S / 16

    def phase_a(self):
        flags = self.eng.sweep_phase(PH_PROPOSE_CONN)
        self.eng.synchronize()
        return flags

    def exchange_a(self, flags):
        if flags & 2:
            self.reduce_fn(self.Sp)
        if flags & 1:
            self.reduce_fn(self.S)

    def phase_b(self):
        self.eng.sweep_phase(PH_DECIDE_Z)
        self.eng.sweep_phase(PH_SWEEP_Y)
        self.eng.synchronize()
        self.S[self.not_own] = 0
        self.y[self.not_own] = 0

    def exchange_b(self):
        self.reduce_fn(self.S)
        self.reduce_fn(self.y)

    def phase_c(self):
        self.torch.cuda.synchronize()
        self.eng.sweep_phase(PH_FINISH)

    def sweep(self, nsweeps=1):
        """One sweep = three compute phases with a collective after the first two.  `self.phase_s` accumulates the host
        time of each stage (every stage ends with a device synchronisation)."""
        import time
        acc = self.__dict__.setdefault("phase_s", dict(conn=0.0, exchange_S=0.0, scan=0.0, exchange_y=0.0, finish=0.0))
        for _ in range(nsweeps):
            t0 = time.perf_counter()
            flags = self.phase_a()
            t1 = time.perf_counter()
            self.exchange_a(flags)
            self.torch.cuda.synchronize()
            t2 = time.perf_counter()
            self.phase_b()
            self.torch.cuda.synchronize()
            t3 = time.perf_counter()
            self.exchange_b()
            self.torch.cuda.synchronize()
            t4 = time.perf_counter()
            self.phase_c()
            self.eng.synchronize()
            t5 = time.perf_counter()
            for k, v in zip(("conn", "exchange_S", "scan", "exchange_y", "finish"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                acc[k] += v
        self.eng.synchronize()


def comm_init_from_torch(eng, group=None):
    """Give the engine an NCCL communicator over the ranks of the torch.distributed job: rank 0 draws the unique id
    (mp_comm_unique_id), torch.distributed only carries its 128 bytes, the communicator itself lives inside the library."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    eng.comm_init(world, rank, box[0])
    return rank, world


class NativeShardedChain:
    """ShardedChain with the collectives inside the library (mp_sweep_sharded): all-gather of the owned connectivity
    columns, broadcast of the owned year rows, no host synchronisation inside a sweep."""

    def __init__(self, eng, group=None):
        self.eng = eng
        self.rank, self.world = comm_init_from_torch(eng, group)

    def describe(self):
        return (f"one chain over {self.world} GPUs: connectivity by target patches, y scan by years; all-gather of the owned "
                "columns of S / S_prop and broadcast of the owned rows of y and S (NCCL inside libmidaspom_cuda.so, on the engine stream)")

    def bytes_per_sweep(self):
        C, T, N = self.eng.C, self.eng.T, self.eng.N
        rows = C * (T - 1)
        return rows * N * 8 * (1 + 1 / 16) + rows * N * 9          # S_prop columns every sweep (+ S every 16th); rows of S and y after the scan

    def sweep(self, nsweeps=1):
        self.eng.sweep_sharded(nsweeps)
        self.eng.synchronize()


def sweep_emulated_ranks(chains, nsweeps=1):
    """Run W ShardedChain objects that live in ONE process (one GPU) in lock step, summing their buffers
    directly instead of through NCCL -- the single-GPU emulation of the multi-rank path used by the tests."""
    def sum_over(name):
        tot = sum(getattr(c, name).to(chains[0].torch.float64) for c in chains)
        for c in chains:
            getattr(c, name).copy_(tot.to(getattr(c, name).dtype))
    for _ in range(nsweeps):
        flags = [c.phase_a() for c in chains]
        chains[0].torch.cuda.synchronize()
        if flags[0] & 2:
            sum_over("Sp")
        if flags[0] & 1:
            sum_over("S")
        chains[0].torch.cuda.synchronize()
        for c in chains:
            c.phase_b()
        chains[0].torch.cuda.synchronize()
        sum_over("S"); sum_over("y")
        for c in chains:
            c.phase_c()
    for c in chains:
        c.eng.synchronize()
