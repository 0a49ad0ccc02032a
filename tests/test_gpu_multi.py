"""Several GPUs through the C ABI (mp_comm_*, mp_gather_draws, mp_sweep_sharded): NCCL inside libmidaspom_cuda.so.
Skipped on boxes with one GPU.  Replaces the MPI plumbing of main_MIDASPOM_MPI.c:344-372,483-505."""
import numpy as np
import pytest

import midaspom_b200 as mb
import oracle_lib as O
from gpu_util import pdict, random_landscape

pytestmark = pytest.mark.gpu


def ngpu():
    return mb.load_library().mp_device_count()


def fresh_engine(spec, device, C, nsw, par, kw, chain_offset=0, blocks=None):
    obs = spec["obs"]
    T, n = obs.shape
    eng = mb.Engine(n, T, C, precision=mb.FP32, device=device, seed=17, max_draws=nsw, chain_offset=chain_offset)
    eng.set_landscape_coords(spec["px"], spec["py"], spec.get("area")); eng.set_source_units(None)
    eng.set_observations(obs)
    if blocks:
        eng.set_scan_blocks(*blocks)
    eng.set_params([par] * C)
    eng.init_chains(mb.engine.sampler_config(**kw), disperse=False)
    return eng


PAR = pdict(e=0.4, c=0.01, alpha=1 / 400, b=0.5)
KW = dict(sample_alpha=1, sample_b=1, c_max=0.2, alpha_min=1e-4, alpha_max=1e-1, n_adapt=4)


@pytest.mark.skipif(ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n", [3000, 5200])
def test_sharded_sweep_over_two_gpus_in_one_process_equals_single_engine(n):
    """mp_comm_init_all + mp_sweep_sharded_all: one set of chains replicated on 2 GPUs, connectivity split by target patches,
    scan by years, NCCL all-gather / broadcast on the engine streams -- bit for bit the single-engine run."""
    rng = np.random.default_rng(31)
    T, C, nsw = 7, 2, 6
    spec, z, y = random_landscape(rng, n, T, O.GEOM_COORDS, occ=0.5, miss=0.05)
    ref = fresh_engine(spec, 0, C, nsw, PAR, KW)
    ref.sweep(nsw)
    want = (ref.get_draws(), ref.get_state(), ref.get_connectivity())
    ref.close()
    engs = [fresh_engine(spec, d, C, nsw, PAR, KW) for d in range(2)]
    mb.Engine.comm_init_all(engs)
    mb.Engine.sweep_sharded_all(engs, nsw)
    for e in engs:
        e.synchronize()
        got = (e.get_draws(), e.get_state(), e.get_connectivity())
        assert (got[0] == want[0]).all()
        assert (got[1][0] == want[1][0]).all() and (got[1][1] == want[1][1]).all()
        assert (got[2] == want[2]).all()
    for e in engs:
        e.close()


@pytest.mark.skipif(ngpu() < 2, reason="needs 2 GPUs")
def test_sharded_sweep_with_scan_blocks_over_two_gpus_equals_single_engine():
    """The same with the block grid of the scan on (21,025 patches, 4 x 4 cells, 9 colours): every rank scans the blocks of its
    years concurrently; the per-year block decisions depend on the year's own data, so nothing changes bit for bit."""
    from test_gpu_blocks import ALPHA, auto_grid, wide_landscape
    rng = np.random.default_rng(33)
    T, C, nsw = 5, 1, 3
    spec, z, y = wide_landscape(rng, 145, T)
    par = pdict(e=0.4, c=0.05, alpha=ALPHA, b=0.5)
    kw = dict(sample_alpha=1, sample_b=1, c_max=0.5, alpha_min=1e-3, alpha_max=1e-1, n_adapt=4)
    grid = auto_grid(spec, z, y, par, 3)
    assert grid[0] * grid[1] >= 16
    ref = fresh_engine(spec, 0, C, nsw, par, kw, blocks=grid)
    ref.work_counters(reset=True)
    ref.sweep(nsw)
    assert ref.work_counters()["scan_blocks"] > 0
    want = (ref.get_draws(), ref.get_state(), ref.get_connectivity())
    ref.close()
    engs = [fresh_engine(spec, d, C, nsw, par, kw, blocks=grid) for d in range(2)]
    mb.Engine.comm_init_all(engs)
    mb.Engine.sweep_sharded_all(engs, nsw)
    for e in engs:
        e.synchronize()
        assert e.work_counters()["scan_blocks"] > 0
        got = (e.get_draws(), e.get_state(), e.get_connectivity())
        assert (got[0] == want[0]).all()
        assert (got[1][0] == want[1][0]).all() and (got[1][1] == want[1][1]).all()
        assert (got[2] == want[2]).all()
    for e in engs:
        e.close()


def _rank_main(rank, world, uid_q, out_q, spec, nsw):
    import midaspom_b200 as mb2
    if rank == 0:
        uid = mb2.Engine.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get()
    # independent chains per rank + mp_gather_draws
    C = 2
    eng = fresh_engine(spec, rank, C, nsw, PAR, KW, chain_offset=rank * C)
    eng.comm_init(world, rank, uid)
    eng.sweep(nsw); eng.synchronize()
    mine = eng.get_draws()
    allr = eng.gather_draws()
    assert allr.shape == (world, nsw, C, mb2.NDRAW) and (allr[rank] == mine).all()
    out_q.put((rank, allr))
    eng.close()


@pytest.mark.skipif(ngpu() < 2, reason="needs 2 GPUs")
def test_gather_draws_over_two_processes():
    """One process per GPU, unique id shipped through a queue (the launcher's job): mp_comm_init + mp_gather_draws give every
    rank all chains; chain c of rank r is global chain r*C + c (its Philox stream), so rank 1's draws equal those of a
    single engine holding chains 2..3."""
    import torch.multiprocessing as tmp
    rng = np.random.default_rng(5)
    spec, z, y = random_landscape(rng, 600, 5, O.GEOM_COORDS, occ=0.5, miss=0.05)
    nsw, world = 5, 2
    ctx = tmp.get_context("spawn")
    uid_q, out_q = ctx.SimpleQueue(), ctx.SimpleQueue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, uid_q, out_q, spec, nsw)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(out_q.get() for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert (res[0] == res[1]).all()
    solo = fresh_engine(spec, 0, 2, nsw, PAR, KW, chain_offset=2)
    solo.sweep(nsw)
    assert (solo.get_draws() == res[0][1]).all()
    solo.close()


@pytest.mark.skipif(ngpu() < 2, reason="needs 2 GPUs")
def test_driver_G_flag_writes_the_draws_of_a_single_gpu_run(tmp_path):
    """driver/midaspom -G 0,1 (the drop-in for `mpirun -np 2 MIDASPOM_MPI.out`, run_examples_MPI.sh:8): chains split over two
    devices, draws gathered by NCCL inside the library (mp_comm_init_all + mp_gather_draws_all).  Philox streams are keyed by
    the global chain id, so the draw file equals that of the same chains on one GPU, line for line."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    subprocess.run(["make", "-C", str(root / "driver")], check=True, capture_output=True)
    example = root / "tests" / "golden" / "occupancies_example.txt"
    outs = []
    for tag, dev in (("one", ["-g", "0"]), ("two", ["-G", "0,1"])):
        res = subprocess.run([str(root / "driver" / "midaspom"), "-m", "400", "-d", "100", "-s", "21", "-n", "600", "-c", "8", "-i", str(example),
                              "-o", str(tmp_path / f"p_{tag}.txt"), "-t", str(tmp_path / f"d_{tag}.txt")] + dev, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        outs.append(((tmp_path / f"d_{tag}.txt").read_text(), (tmp_path / f"p_{tag}.txt").read_text(), res.stdout))
    assert "on 2 GPUs" in outs[1][2]
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
