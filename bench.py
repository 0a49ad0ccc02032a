#!/usr/bin/env python
"""bench.py -- MCMC iterations/s of the SPOM engine on the shapes BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|tiny] [--impl ours|reference]

A "step" is one MCMC iteration (sweep) of every chain resident on the GPU: refresh of the
connectivity S, Metropolis steps on (alpha, b) and c, Gibbs updates of the latent z cells and of
every free intermediate-state cell y (rank-1 updates of S), Metropolis steps on e (and p).
metric = chain-iterations per second = (chains on all ranks) x K / (max over ranks of the device
time of the K steps).  Default workload: cfg3 = synthetic N=10,000 patches x T=20 years, 8 chains
per GPU (64 chains on 8 GPUs, weak scaling).  Under torchrun each rank owns its own block of
chains (no data-path collective); NCCL is used only to gather the draws for R-hat.

--impl reference times the CPU side (the reference's algorithm cannot run at N >= 32, so this is
the FP64 C restatement of its per-cell terms in oracle/, one chain per host core, on a bounded
sample of the same sweep) and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "mcmc_chain_iterations_per_sec"
UNIT = "chain-iterations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ess-sweeps", type=int, default=1000, help="extra untimed-by-step sweeps after the timed regions for the ESS/s figure (0 = skip)")
    return ap.parse_args()


def sampler_kwargs(wl):
    c0 = wl["truth"]["c"]
    return dict(sample_e=1, sample_c=1, sample_alpha=1, sample_b=1, sample_p=int(wl["detect"]),
                e_min=0.0, e_max=1.0, c_min=0.0, c_max=20.0 * c0, alpha_min=1e-4, alpha_max=1e-1, b_min=0.0, b_max=2.0,
                p_min=0.0, p_max=1.0, n_e_steps=4, n_c_steps=1, n_adapt=200, update_z=1, update_y=1,
                sample_K=int("era" in wl and "src_unit" not in wl), K_min=0.1, K_max=100.0, n_v_steps=2,   # die-off variant: K on the reference's range (dieoff.c:113-114)
                sample_Ksrc=int("src_unit" in wl), sample_dsrc=int("src_unit" in wl),                       # patch-loss variant: K_L, d_L (loss.c:93-101)
                Ksrc_min=0.1, Ksrc_max=100.0, dsrc_min=100.0, dsrc_max=4000.0)


def start_params(wl):
    t = wl["truth"]
    return dict(e=0.5 if "era" not in wl else t["e"], c=t["c"], alpha=t["alpha"], b=t["b"], p=t["p"], K=t.get("K", 1.0), Ksrc=t.get("Ksrc", 0.0), dsrc=t.get("dsrc", 0.0))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ----------------------------------------------------------------------------- CPU side (oracle port)
def cpu_sample(wl, chains, budget_s, steps=1, warmup=0):
    """Time the oracle's sweep (FP64 C restatement, one chain per core) on a bounded sample:
    the fixed part of a sweep (two full connectivity evaluations + parameter updates) in full,
    the y scan on `limit` candidate cells per year, extrapolated to all candidates."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    m = O.Model(wl["obs"], geom=O.GEOM_COORDS, px=wl["px"], py=wl["py"], area=wl["area"], detect=wl["detect"], era=wl.get("era"))
    cfg = O.sampler_cfg(**sampler_kwargs(wl))
    sp = start_params(wl)
    cores = O.lib().spom_max_threads()
    nch = max(1, min(chains, cores))
    ch = O.Chains(m, cfg, nch, seed=1000, par0=O.params(**sp), disperse=False)
    ncand = int(((ch.z[0][:-1] & ch.z[0][1:]) == 1).sum())          # candidate cells per chain-sweep
    # probe with 1 candidate per year to size the sample, then the measured sweeps
    ch.run(1, y_flip_limit=1, nthreads=cores)
    t_fixed0 = float(ch.phase_s[:, 0].max())
    per_flip = max(float(ch.phase_s[:, 1].max()) / max(1.0, ch.visited / nch), 1e-7)   # s per candidate (one chain, one core)
    per_sweep_budget = budget_s / max(1, steps + warmup)
    per_year = int((per_sweep_budget - t_fixed0) / (per_flip * (wl["T"] - 1))) if per_sweep_budget > t_fixed0 else 1
    per_year = int(min(max(per_year, 1), wl["n"]))
    fixed, yscan, visited = [], [], []
    for s in range(warmup + steps):
        ch.run(1, y_flip_limit=per_year, nthreads=cores)
        if s >= warmup:
            fixed.append(float(ch.phase_s[:, 0].max())); yscan.append(float(ch.phase_s[:, 1].max())); visited.append(ch.visited / nch)
    t_fixed, t_y, vis = float(np.mean(fixed)), float(np.mean(yscan)), float(np.mean(visited))
    t_step = t_fixed + t_y
    t_full = t_fixed + t_y * (ncand / max(vis, 1.0))                  # one sweep of nch chains, one per thread
    value = nch / t_full                                              # chain-iterations/s with all cores busy
    sample = (f"{nch} chains on {cores} threads, FP64 oracle port (CPU restatement of the reference's per-cell terms, "
              f"not MIDASPOM_MPI.out: the reference enumerates 2^N states); per sweep: full connectivity refresh + "
              f"proposal evaluated in full ({t_fixed:.2f} s), y scan timed on {int(vis)} of {ncand} candidate cells per chain "
              f"({t_y:.3f} s) and extrapolated linearly to all of them")
    return dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample, t_fixed_s=t_fixed, t_step_s=t_step,
                ms_per_step=t_full * 1e3, measured_steps=len(fixed))


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload.startswith("cfg5"):
        emit(json.dumps(dict(impl="reference", unavailable="N=100,000: one connectivity evaluation is 1e10 pair terms (minutes per "
                              "sweep on the host); no CPU arm for this workload")), flush=True)
        return
    from midaspom_b200 import synth
    wl = synth.make_workload(args.workload)
    chains = wl["chains_per_gpu"] * max(1, args.gpus)
    r = cpu_sample(wl, chains, args.cpu_seconds * 3, steps=args.steps, warmup=min(args.warmup, 1))
    line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=f"{args.workload}: {wl['desc']}", n_patches=wl["n"], n_years=wl["T"], chains=chains,
                            geometry="planar coordinates + areas"),
                cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import midaspom_b200 as mb
    from midaspom_b200 import synth, distributed as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = synth.make_workload(args.workload)
    cpg = wl["chains_per_gpu"]
    if args.workload.startswith("cfg5") and world > 1:
        return run_sharded(args, rank, local_rank, world, wl)
    first, _ = D.chain_block(rank, world, cpg)
    K, W = args.steps, args.warmup
    n, T = wl["n"], wl["T"]
    eng = mb.Engine(n, T, cpg, precision=mb.FP32, device=local_rank, seed=1000, detect=wl["detect"], chain_offset=first,
                    max_draws=2 * K + W + 8 + max(0, args.ess_sweeps))
    eng.set_landscape_coords(wl["px"], wl["py"], wl["area"])
    eng.set_source_units(wl.get("src_unit"))
    obs_pinned = torch.from_numpy(wl["obs"].copy()).pin_memory()
    obs_host = obs_pinned.numpy()
    eng.set_observations(obs_host)
    eng.set_era(wl.get("era"))
    eng.set_params([start_params(wl)] * cpg)
    eng.init_chains(mb.engine.sampler_config(**sampler_kwargs(wl)), disperse=False)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(nsteps, e2e=False):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        draws = np.zeros((1, cpg, mb.NDRAW))
        with torch.cuda.stream(stream):
            for a, b in evs:
                flush.fill_(1)                                   # evict the chain state from L2 between steps
                a.record(stream)
                if e2e:
                    eng.set_observations(obs_host)               # H2D from pinned memory, every step
                    eng.sweep(1, sync=False)
                    draws = eng.get_draws(eng.num_draws() - 1, 1)  # D2H of the step's result (syncs)
                else:
                    eng.sweep(1, sync=False)
                b.record(stream)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], draws

    eng.sweep(W)                                                 # warm-up (untimed)
    barrier()
    eng.set_timing(True)
    eng.get_timing(reset=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        step_ms, _ = timed_steps(K)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    kms, klaunch = eng.get_timing(reset=True)
    eng.set_timing(False)
    total_ms = float(sum(step_ms))
    # end-to-end arm: host buffers in and out every step
    barrier()
    e2e_ms, last = timed_steps(K, e2e=True)
    barrier()
    e2e_total = float(sum(e2e_ms))
    t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_total = float(t[0]), float(t[1])
    chains_total = cpg * world
    value = chains_total * K / (total_ms * 1e-3)
    e2e_value = chains_total * K / (e2e_total * 1e-3)

    # ESS/s (third part of BASELINE's metric): a longer run after the timed regions, wall clock, second half of its draws
    ess_run = None
    n_ess = min(args.ess_sweeps, int(15000.0 / max(total_ms / K, 1e-3)))      # at most ~15 s of extra sweeps
    if n_ess >= 200:
        barrier()
        t0 = time.perf_counter()
        eng.sweep(n_ess)
        barrier()
        ess_run = time.perf_counter() - t0
    # likelihood evaluations/s (second part of BASELINE's metric): 1 evaluation = full connectivity S of the resident
    # state + the summed log-terms of one chain, through the public calls mp_connectivity + mp_loglik
    barrier()
    n_ll = 20
    t0 = time.perf_counter()
    for _ in range(n_ll):
        eng.connectivity(fetch=False)
        ll_last, _parts = eng.loglik()
    barrier()
    t_ll = time.perf_counter() - t0
    tl = torch.tensor([t_ll], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    lik_evals = chains_total * n_ll / float(tl[0])
    # posterior diagnostics on everything recorded so far (gathered over ranks with NCCL)
    nd = eng.num_draws()
    d_local = torch.from_numpy(eng.get_draws(0, nd)).to(f"cuda:{local_rank}")
    d_all = D.gather_draws(d_local).cpu().numpy()
    if ess_run is not None:
        half = n_ess // 2
        summ = D.posterior_summary(d_all[nd - half:])
        run_s = ess_run * half / n_ess
    else:
        summ = D.posterior_summary(d_all[W:])
        run_s = (total_ms + e2e_total) * 1e-3
    ess_min = min((v["ess"] for v in summ.values()), default=float("nan"))

    if rank == 0:
        z, y = eng.get_state()
        ncand = float(((z[:, :-1] & z[:, 1:]) == 1).sum()) / cpg             # candidate cells per chain-sweep
        peaks_meas = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
        hbm_peak = float(peaks_meas.get("hbm_gbs", 6650.0))
        probe = eng.probe_peaks()
        n_sweepy = max(1, klaunch["sweep_y"])
        t_sweepy = kms["sweep_y"] / n_sweepy * 1e-3                           # s per launch (CUDA events, engine stream)
        pairs = cpg * ncand * (n - 1)                                         # pair evaluations per launch
        mufu_per_pair = 2.2                                                   # MUFU.SQRT + MUFU.EX2 + MUFU.LG2 / 5 (planar geometry)
        alg_bytes = cpg * (T - 1) * n * (8 + 1 + 1 + 1 + 8 + 1) + cpg * ncand * 32   # S,y,z_t,z_t+1 in; S,y out; candidate records
        t_conn = kms["conn"] / max(1, klaunch["conn"]) * 1e-3
        conn_pairs = cpg * float(n) * n * (klaunch['conn'] and (1.0 + 1.0 / 16.0))   # proposal every sweep + resident S every 16th
        traffic = None                                                        # dram read+write per launch, from the tracked ncu summary
        for f in sorted((ROOT / "profiles").glob("*_ncu_full.json"), reverse=True):
            for kd in json.loads(f.read_text()):
                if "k_sweep_y" in kd.get("kernel", "") and "dram_traffic_bytes_per_launch" in kd:
                    traffic = dict(bytes_per_launch=kd["dram_traffic_bytes_per_launch"], source=f"profiles/{f.name}")
                    break
            if traffic:
                break
        sweep_kernel = "k_sweep_y_cull" if n > 3000 else "k_sweep_y_fast"     # planar landscapes above 3,000 patches take the culled scan
        roof = dict(bound="sfu", kernel=sweep_kernel, achieved=pairs * mufu_per_pair / t_sweepy * 1e-9,
                    peak=probe["mufu_gops"], unit="Gop/s (MUFU)", frac=pairs * mufu_per_pair / t_sweepy * 1e-9 / probe["mufu_gops"],
                    traffic=traffic, share_of_step=kms["sweep_y"] / max(total_ms, 1e-9), ms_per_launch=t_sweepy * 1e3,
                    algorithmic=dict(pairs_per_launch=pairs, mufu_per_pair=mufu_per_pair, bytes_per_launch=alg_bytes,
                                     note="pairs = candidates x (N-1) targets, the dense algorithm's count; the culled scan evaluates "
                                          "only the targets within FP32 reach of a candidate (cfg3: 25.5% of them, measured) and "
                                          "skips the rest exactly, so 'achieved' counts work the kernel avoids as done"),
                    peak_source="mp_probe_peaks micro-benchmark on this GPU (MEASURED_PEAKS.json has no MUFU figure)",
                    hbm=dict(achieved=alg_bytes / t_sweepy * 1e-9, peak=hbm_peak, unit="GB/s", frac=alg_bytes / t_sweepy * 1e-9 / hbm_peak,
                             peak_source="MEASURED_PEAKS.json" if peaks_meas else "fallback"),
                    conn=dict(kernel="k_conn", ms_per_launch=t_conn * 1e3, achieved=conn_pairs * 2 / t_conn * 1e-9,
                              unit="Gop/s (MUFU)", frac=conn_pairs * 2 / t_conn * 1e-9 / probe["mufu_gops"],
                              share_of_step=kms["conn"] / max(total_ms, 1e-9)),
                    probe=probe)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=total_ms / K,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=f"{args.workload}: {wl['desc']}", n_patches=n, n_years=T, chains=chains_total,
                                chains_per_gpu=cpg, geometry="planar coordinates + areas, on-the-fly weights",
                                sampled=["e", "c", "alpha", "b"] + (["p"] if wl["detect"] else []) + (["Ksrc", "dsrc"] if "src_unit" in wl else ["K"] if "era" in wl else []),
                                l2="flushed between timed steps (256 MiB write)", parallelism=f"chains x{world}"),
                    clocks=clk.summary(),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(obs_host.nbytes),
                             d2h_bytes_per_step=int(cpg * mb.NDRAW * 8), ms_per_step=e2e_total / K),
                    gpu_launches=int(sum(klaunch.values())),
                    kernel_ms={k: round(v, 4) for k, v in kms.items()}, kernel_launches=klaunch,
                    wall_s_timed_region=t_wall,
                    likelihood_evals_per_sec=lik_evals,
                    ess_per_sec=ess_min / run_s if run_s > 0 else None,
                    ess=dict(min_ess=ess_min, seconds=run_s, sweeps=(n_ess // 2 if ess_run is not None else nd - W),
                             note="min over sampled parameters of the summed per-chain ESS (Geyer), second half of a separate "
                                  f"{n_ess}-sweep run, wall clock" if ess_run is not None else "timed draws only"),
                    posterior=summ, candidates_per_chain_sweep=ncand, roofline=roof)
        if not args.no_cpu_baseline and world == 1:
            try:
                r = cpu_sample(wl, cpg, args.cpu_seconds, steps=1, warmup=0)
                line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"])
            except Exception as ex:  # the oracle is test infrastructure: never let it break the GPU line
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=0, kind="port", sample=f"failed: {ex}")
        emit(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- cfg5 on several GPUs: one chain, sharded
def run_sharded(args, rank, local_rank, world, wl):
    """BASELINE config 5: one chain at N=100,000 x T=30 sharded over the GPUs of the box -- connectivity by target
    patches, y scan by years, an all-reduce (NCCL) of S / S_prop and of the occupancy state y per sweep
    (midaspom_b200/distributed.py: ShardedChain).  Total work is fixed: "scaling": "strong"."""
    import torch
    import torch.distributed as dist
    import midaspom_b200 as mb
    from midaspom_b200 import distributed as D
    K, W = args.steps, args.warmup
    n, T, C = wl["n"], wl["T"], wl["chains_per_gpu"]
    dev = torch.device("cuda", local_rank)
    eng = mb.Engine(n, T, C, precision=mb.FP32, device=local_rank, seed=1000, detect=wl["detect"], chain_offset=0, max_draws=K + W + 4)
    eng.set_landscape_coords(wl["px"], wl["py"], wl["area"]); eng.set_source_units(None)
    eng.set_observations(wl["obs"])
    eng.set_params([start_params(wl)] * C)
    eng.init_chains(mb.engine.sampler_config(**sampler_kwargs(wl)), disperse=False)
    sc = D.ShardedChain(eng, rank, world, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sc.sweep(W)
    sc.phase_s = dict.fromkeys(sc.phase_s, 0.0)
    times = []
    with ClockSampler(local_rank) as clk:
        for _ in range(K):
            flush.fill_(1)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            sc.sweep(1)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    t = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_s = float(t[0])
    draws = eng.get_draws()
    same = torch.tensor(draws[-1, 0, :6], device=dev)
    ref = same.clone(); dist.broadcast(ref, 0)
    identical = bool((same == ref).all())
    ok = torch.tensor([1.0 if identical else 0.0], device=dev); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        kms, klaunch = eng.get_timing(reset=False)
        bytes_per_sweep = 2 * C * (T - 1) * n * 8 + C * (T - 1) * n + (C * (T - 1) * n * 8) / 16   # S + y every sweep, S_prop; refresh /16
        line = dict(metric=METRIC, value=C * K / total_s, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=total_s / K * 1e3,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=f"{args.workload}: {wl['desc']}", n_patches=n, n_years=T, chains=C,
                                parallelism=f"one chain over {world} GPUs: connectivity by target patches, y scan by years, "
                                            "all-reduce of S and y per sweep (NCCL)", l2="flushed between timed steps (256 MiB write)",
                                timing="host clock between device synchronisations, max over ranks"),
                    clocks=clk.summary(), gpu_launches=int(sum(klaunch.values())),
                    collective_bytes_per_sweep=int(bytes_per_sweep), ranks_hold_identical_draws=bool(ok.item() == 1.0),
                    phase_ms_per_sweep={k: round(v / K * 1e3, 3) for k, v in sc.phase_s.items()},
                    e2e=dict(value=C * K / total_s, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0,
                             note="state resident; the sharded loop has no per-step host buffers"))
        emit(json.dumps(line), flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


# ----------------------------------------------------------------------------- cfg1: the bundled example, exact grid posterior
def run_cfg1(args, rank):
    """BASELINE config[0]: `MIDASPOM.out -m 400 -d 100` on the bundled 8 x 7 example = 101 x 101 = 10,201 exact
    likelihood evaluations (run_examples.sh:8).  ours: mp_exact_posterior (host buffers in and out);
    reference: the reference binary itself, compiled from its sources into oracle/_ref (kind "reference")."""
    if rank != 0:
        return
    import tempfile
    example = ROOT / "tests" / "golden" / "occupancies_example.txt"
    nev = 101 * 101
    cfgd = dict(workload="cfg1: bundled example (8 patches x 7 years), 101 x 101 grid of (e, c), exact likelihood", grid=101)
    if args.impl == "reference":
        exe = ROOT / "oracle" / "_ref" / "MIDASPOM.out"
        if not exe.exists():
            emit(json.dumps(dict(impl="reference", unavailable="oracle/_ref/MIDASPOM.out not built (no /root/reference at build time)")))
            return
        tmp = tempfile.mkdtemp()
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            subprocess.run([str(exe), "-m", "400", "-d", "100", "-i", str(example), "-o", f"{tmp}/p.txt"], check=True, capture_output=True)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = float(np.mean(times))
        v = nev / t
        emit(json.dumps(dict(metric="likelihood_evals_per_sec", value=v, unit="likelihood evaluations/s", n_gpus=args.gpus, steps=args.steps,
                              warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                              data="bundled example", impl="reference", config=cfgd,
                              cpu_baseline=dict(value=v, unit="likelihood evaluations/s", cores=1, kind="reference",
                                                sample="MIDASPOM.out (reference sources, gcc -O3, naive cblas_dgemm), whole program incl. file I/O, single thread"),
                              e2e=dict(value=v, unit="likelihood evaluations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)))
        return
    import torch
    import midaspom_b200 as mb
    raw = example.read_bytes()
    n = 1 + sum(1 for ch in raw.split(b"\n", 1)[0] if ch in b" \t")
    tmax = raw.count(b"\n")
    obs = np.array([int(v) for v in raw.split()][: n * tmax], dtype=np.int8).reshape(tmax, n)
    times = []
    for s in range(args.warmup + args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll, ltot, info = mb.exact_posterior(obs, a=1 / 400, d=100.0, prior_occ=0.5, nstep=101)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    v = nev / t
    emit(json.dumps(dict(metric="likelihood_evals_per_sec", value=v, unit="likelihood evaluations/s", n_gpus=1, steps=args.steps, warmup=args.warmup,
                          ms_per_step=t * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="bundled example",
                          config=dict(cfgd, states=info["nstates"], short_states=info["nextid"], total_loglik=ltot),
                          ms_per_step_min=float(np.min(times)) * 1e3, ms_per_step_median=float(np.median(times)) * 1e3,
                          e2e=dict(value=v, unit="likelihood evaluations/s", h2d_bytes_per_step=int(obs.nbytes), d2h_bytes_per_step=nev * 8),
                          gpu_launches=2 * args.steps)))


_REAL_STDOUT = None


def emit(line, flush=True):
    """The ONE JSON line of the contract goes to the process's real stdout; everything else any library writes to
    file descriptor 1 (NCCL prints its version there under torchrun) has been redirected to stderr by main()."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(line + "\n")
    if flush:
        out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.workload == "cfg1":
        run_cfg1(args, rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
