#!/usr/bin/env python3
"""Throughput of the device-resident PTMCMC step on a BASELINE ensemble: chain-steps per second (one
chain-step = proposal + prior + likelihood + accept), against the likelihood-only rate of the same ensemble through the C ABI
(the ceiling) and the reference's CPU likelihood rate (the ceiling of the reference's own sampler).  One JSON line per run.

    python tools/bench_sampler.py [--config 2] [--steps 200] [--lanes 2] [--no-fisher]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gw_analysis_tools_b200 import engine, workloads  # noqa: E402
from gw_analysis_tools_b200 import sampler as smp  # noqa: E402

N_TEMPS = {1: 8, 2: 8, 4: 16, 5: 8}


def run_under_bench(args, ctx, world, rank, local, dist, ClockSampler):
    """`bench.py --workload sampler [--gpus N] [--scaling strong|weak]`: device-resident PTMCMC steps on the config's ensemble.

    strong (what north_star describes): the config's ONE ladder (cfg2: 8 temperatures x 512 ensembles = 4096 chains) is cut into
    N contiguous blocks of chains, one per rank; every swap sweep all-gathers (position, logL, logP) over NCCL inside the library
    (gwat_b200_sampler_attach_ranks) and every rank runs the reference's sweep over the whole ladder.  weak: every rank owns a
    ladder of the config's size and joins the same exchange (N x 4096 chains).  One "step" = one MH step of every chain.
    value = chain-steps/s from the device time of gwat_b200_sampler_run (CUDA events on the sampler's stream), max over ranks;
    e2e = the same from wall clock around the call, including a device->host read of the cold chains' state every `steps`."""
    import torch
    cfg = args.config if args.config in N_TEMPS else 2
    wl = workloads.make(cfg, W=args.walkers, L=args.bins)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    wl.data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)
    nt = N_TEMPS[cfg]
    strong = args.scaling == "strong"
    C_total = wl.W if strong else wl.W * world
    C_loc = C_total // world
    ladder = np.tile(np.geomspace(1.0, 100.0, nt), C_total // nt)
    rng = np.random.default_rng(11)
    reps = -(-C_total // wl.W)
    base = np.concatenate([wl.params] * reps)[:C_total]
    init_all = wl.inj[None, :] + 0.2 * (base - wl.inj[None, :])
    if reps > 1:
        init_all[:, [0, 2, 4]] += 1e-3 * rng.standard_normal((C_total, 3))
    lo = rank * C_loc
    prior = smp.prior_for(wl)
    fisher = 1
    s = smp.Sampler(ctx, wl.method, ladder[lo:lo + C_loc], init_all[lo:lo + C_loc], prior, wl.gmst, wl.T_segment, wl.mod, seed=1, lanes=2,
                    fisher_exist=fisher, fisher_update_number=200, history_length=1000, fisher_deferred=1, chain_index_offset=lo)
    if world > 1:
        ids = [smp.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        s.attach_ranks(ids[0], rank, world)
    s.run(max(args.warmup, 10))
    # the clock sampler starts BEFORE the barrier: NVML start-up takes 10-20 ms and differs between ranks, and a rank that enters
    # the timed call late makes every other rank wait for it at the first swap exchange (seen as 1.5-2 ms "per sweep" on 8 GPUs)
    sampler = ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(args.steps)
    pos, ll, lp = s.state()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms, swap_ms, sweeps, launches = s.last_ms, s.last_swap_ms, s.last_sweeps, s.last_launches
    if world > 1:
        t = torch.tensor([dev_ms, wall, swap_ms], dtype=torch.float64, device="cuda")
        allv = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per_rank = [[float(x) for x in a] for a in allv]
        dev_ms, wall, swap_ms = (max(col) for col in zip(*per_rank))
        swap_ms_min = min(r[2] for r in per_rank)
    else:
        per_rank = [[dev_ms, wall, swap_ms]]
        swap_ms_min = swap_ms
    ct, _ = s.counters()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    P = wl.P
    line = {"metric": "PTMCMC chain-steps/sec (proposal + prior + waveform/response/inner-product likelihood + MH accept; PT swap sweep every %d steps)" % s.options.swp_freq,
            "value": C_total * args.steps / (dev_ms * 1e-3), "unit": "chain-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 10),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s sampler: %s, %d detectors, %d chains in all (%d ensembles x %d temperatures), %d per GPU, %d bins, Gaussian + differential-evolution + Fisher proposals" % (
                           wl.name, wl.method, wl.D, C_total, C_total // nt, nt, C_loc, wl.L),
                       "method": wl.method, "chains_total": C_total, "chains_per_gpu": C_loc, "bins": wl.L, "detectors": wl.D, "dimension": P,
                       "parallelism": ("one ladder sharded over %d GPU(s)" % world if strong else "one ladder of %d x the config's chains over %d GPU(s)" % (world, world)) +
                                      "; swap sweep = ncclAllGather of (position, logL, logP) per chain inside the library" if world > 1 else "single GPU, swap sweep on the device",
                       "l2": "a fresh proposal set every step (the chains move); inputs stay resident: this is the sampler, not a copy benchmark"},
            "e2e": {"value": C_total * args.steps / wall, "unit": "chain-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": C_loc * (P + 2) * 8 / args.steps, "ms_per_step": wall / args.steps * 1e3,
                    "note": "wall clock around gwat_b200_sampler_run + one device->host read of the ensemble state per call"},
            "swap_exchange": {"ms_per_sweep": swap_ms, "ms_per_sweep_last_arriving_rank": swap_ms_min, "sweeps": int(sweeps), "bytes_per_rank_per_sweep": C_loc * (P + 2) * 8,
                              "bytes_gathered_per_sweep": C_total * (P + 2) * 8,
                              "share_of_step_time": swap_ms * sweeps / dev_ms if dev_ms > 0 else None,
                              "what": ("pack + ncclAllGather + thresholds + run/pointer-doubling sweep over the whole ladder + take, CUDA events on the sampler's stream; "
                                       "ms_per_sweep = max over ranks and includes waiting for the slowest rank to reach the collective, the rank that arrives last "
                                       "waits for nobody: its time is the exchange itself") if world > 1 else "thresholds + run/pointer-doubling sweep + counters + moves on the device, no exchange"},
            "per_rank": {"device_ms": [r[0] for r in per_rank], "wall_s": [r[1] for r in per_rank], "swap_ms_per_sweep": [r[2] for r in per_rank]},
            "gpu_launches": int(launches), "clocks": clocks,
            "accept_fraction": float(ct["step_accept"].sum() / max(1, ct["step_accept"].sum() + ct["step_reject"].sum())),
            "swap_accept_fraction": float(ct["swap_accept"].sum() / max(1, ct["swap_accept"].sum() + ct["swap_reject"].sum())),
            "finite": bool(np.isfinite(ll).all())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--lanes", type=int, default=2)
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--bins", type=int, default=0)
    ap.add_argument("--no-fisher", action="store_true")
    ap.add_argument("--fisher-update", type=int, default=200)
    ap.add_argument("--history", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--deferred", type=int, default=0)
    args = ap.parse_args()
    wl = workloads.make(args.config, W=args.chains or None, L=args.bins or None)
    ctx = engine.Context(0)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    wl.data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)
    C_ = wl.W
    nt = N_TEMPS[args.config]
    temps = np.tile(np.geomspace(1.0, 100.0, nt), C_ // nt)
    init = wl.inj[None, :] + 0.2 * (wl.params - wl.inj[None, :])
    prior = smp.prior_for(wl)
    # likelihood-only ceiling on the same ensemble (host buffers in and out, like bench.py's e2e arm)
    for _ in range(3):
        ctx.loglike_mcmc_batch(wl.method, init, wl.gmst, wl.T_segment, wl.mod)
    t0 = time.perf_counter()
    n_like = 20
    for _ in range(n_like):
        ctx.loglike_mcmc_batch(wl.method, init, wl.gmst, wl.T_segment, wl.mod)
    like_rate = n_like * C_ / (time.perf_counter() - t0)
    s = smp.Sampler(ctx, wl.method, temps, init, prior, wl.gmst, wl.T_segment, wl.mod, seed=1, lanes=args.lanes,
                    fisher_exist=0 if args.no_fisher else 1, fisher_update_number=args.fisher_update, history_length=args.history,
                    fisher_deferred=args.deferred)
    s.run(args.warmup)
    t0 = time.perf_counter()
    s.run(args.steps)
    wall = time.perf_counter() - t0
    dev_ms = s.last_ms
    ct, widths = s.counters()
    pos, ll, lp = s.state()
    ctx.set_kernel_timing(True)
    ctx.loglike_mcmc_batch(wl.method, pos, wl.gmst, wl.T_segment, wl.mod)  # where the ensemble is now: active bins, kernel time
    final_active, final_kernel_ms = ctx.last_active_bins / (C_ * wl.L), ctx.last_kernel_ms
    ctx.loglike_mcmc_batch(wl.method, init, wl.gmst, wl.T_segment, wl.mod)
    init_active, init_kernel_ms = ctx.last_active_bins / (C_ * wl.L), ctx.last_kernel_ms
    line = {"metric": "PTMCMC chain-steps/sec (%s, %d chains = %d ensembles x %d temperatures, %d bins, %d detectors)" % (
                wl.method, C_, C_ // nt, nt, wl.L, wl.D),
            "value": C_ * args.steps / (dev_ms * 1e-3), "unit": "chain-steps/s", "wall_value": C_ * args.steps / wall,
            "ms_per_step": dev_ms / args.steps, "steps": args.steps, "lanes": args.lanes, "fisher": not args.no_fisher, "fisher_deferred": args.deferred,
            "likelihood_only_evals_per_s": like_rate, "fraction_of_likelihood_ceiling": C_ * args.steps / wall / like_rate,
            "launches_per_step": s.last_launches / args.steps,
            "accept_fraction": float(ct["step_accept"].sum() / (ct["step_accept"].sum() + ct["step_reject"].sum())),
            "swap_accept_fraction": float(ct["swap_accept"].sum() / max(1, ct["swap_accept"].sum() + ct["swap_reject"].sum())),
            "fisher_updates": int(ct["fisher_updates"].sum()), "fisher_nan": int(ct["fisher_nan"].sum()),
            "finite": bool(np.isfinite(ll).all()),
            "active_bin_fraction": {"initial": init_active, "final": final_active},
            "k_loglike_ms_full_ensemble": {"initial": init_kernel_ms, "final": final_kernel_ms}}
    if args.cpu_sample:
        from oracle import gwat_ref
        if gwat_ref.available():
            n = min(args.cpu_sample, C_)
            threads = max(1, len(os.sched_getaffinity(0)))
            t0 = time.perf_counter()
            gwat_ref.loglike_mcmc_batch(wl.method, wl.mod, init[:n], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data, nthreads=threads)
            line["cpu_reference_likelihood_evals_per_s"] = n / (time.perf_counter() - t0)
            line["cpu_threads"] = threads
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
