"""World-size-2 gloo tests (CPU) of the host logic around chain sharding: the block split, the
all_gather of draws and the R-hat / ESS summary every rank computes from the gathered chains."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from midaspom_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cpr, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = D.chain_block(rank, world, cpr)
    # synthetic "draws": value encodes (sweep, global chain) so the gathered layout can be checked
    sweeps = 64
    g = torch.Generator().manual_seed(1000 + rank)
    draws = torch.zeros(sweeps, count, 8, dtype=torch.float64)
    for c in range(count):
        draws[:, c, 0] = torch.randn(sweeps, generator=g) * 0.1 + 0.5
        draws[:, c, 7] = torch.arange(sweeps) * 1000 + (first + c)
    allc = D.gather_draws(draws)
    summ = D.posterior_summary(allc.numpy(), fields=("e",))
    if rank == 0:
        out.put((tuple(allc.shape), allc[:, :, 7].numpy().copy(), summ))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_and_rhat():
    world, cpr = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cpr, q)) for r in range(world)]
    for p in procs:
        p.start()
    shape, tags, summ = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert shape == (64, world * cpr, 8)
    want = np.arange(64)[:, None] * 1000 + np.arange(world * cpr)[None, :]
    assert (tags == want).all()                     # chain c of rank r lands at global column r*cpr + c
    assert 0.9 < summ["e"]["rhat"] < 1.2 and summ["e"]["ess"] > 100


def test_reference_row_split_covers_every_row_once():
    """split_rows is MIDASPOM_MPI's partition (rank 0 takes the remainder, :361-372)."""
    for total, world in ((101, 4), (151, 4), (7, 3), (8, 8), (5, 1)):
        seen = []
        for r in range(world):
            a, b = D.split_rows(total, world, r)
            seen += list(range(a, b))
        assert seen == list(range(total))
    assert D.split_rows(101, 4, 0) == (0, 26) and D.split_rows(101, 4, 1) == (26, 51)
    assert D.chain_block(3, 8, 8) == (24, 8)
