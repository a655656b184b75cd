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
