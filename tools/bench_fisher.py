#!/usr/bin/env python3
"""Throughput of the batched Fisher path (BASELINE config 3): order-4 numerical Fisher matrices of IMRPhenomD, 11 parameters,
summed over 3 detectors, L bins.  Prints one JSON line: GPU Fisher/s through the host-buffer C ABI, the reference's CPU rate
on a bounded sample (oracle/_ref, all host threads), and the parity measure between the two on that sample.

    python tools/bench_fisher.py [--sources 4096] [--bins 4096] [--cpu-sample 64]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=4096)
    ap.add_argument("--bins", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--method", default="IMRPhenomD")
    ap.add_argument("--dim", type=int, default=11)
    args = ap.parse_args()
    dets = ["Hanford", "Livingston", "Virgo"]
    f = 20.0 + 0.25 * np.arange(args.bins)
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    srcs = workloads.fisher_sources(args.sources)
    ctx = engine.Context(0)
    ctx.set_network(dets, f, psd)
    ctx.fisher_numerical_batch(args.method, srcs, args.dim, order=4)  # warm-up at full size: scratch buffers reach their steady state
    t0 = time.perf_counter()
    F = ctx.fisher_numerical_batch(args.method, srcs, args.dim, order=4)
    dt = time.perf_counter() - t0
    line = {"metric": "order-4 numerical Fisher matrices/sec (%s, dim %d, 3 detectors summed, %d bins)" % (args.method, args.dim, args.bins),
            "value": args.sources / dt, "unit": "Fisher/s", "sources": args.sources, "seconds": dt,
            "device_ms": ctx.last_kernel_ms, "finite": bool(np.all(np.isfinite(F)))}
    bad = np.flatnonzero(~np.all(np.isfinite(F.reshape(args.sources, -1)), axis=1))
    line["nonfinite_sources"] = int(bad.size)
    try:
        from oracle import gwat_ref
        if gwat_ref.available() and bad.size:
            # sources whose Fisher matrix is not finite: does the reference agree?  (checked on up to 32 of them)
            chk = bad[:32]
            R = gwat_ref.fisher_numerical_batch(args.method, [srcs[int(i)] for i in chk], dets, f, psd, args.dim, order=4,
                                                detector_index=-1, reference_index=0)
            line["nonfinite_checked"] = int(chk.size)
            line["nonfinite_in_reference_too"] = int(np.sum(~np.all(np.isfinite(R.reshape(chk.size, -1)), axis=1)))
        if gwat_ref.available() and args.cpu_sample > 0:
            n = min(args.cpu_sample, args.sources)
            threads = max(1, len(os.sched_getaffinity(0)))
            t0 = time.perf_counter()
            R = gwat_ref.fisher_numerical_batch(args.method, srcs[:n], dets, f, psd, args.dim, order=4, detector_index=-1,
                                                reference_index=0, nthreads=threads)
            dtc = time.perf_counter() - t0
            dg = np.sqrt(np.abs(np.einsum("sii->si", R)))
            nerr = np.abs(F[:n] - R) / (dg[:, :, None] * dg[:, None, :])
            line["cpu_baseline"] = {"value": n / dtc, "unit": "Fisher/s", "cores": threads, "kind": "reference",
                                    "sample": "%d sources, %.2f s wall" % (n, dtc)}
            line["parity"] = {"normalised_error_median": float(np.median(nerr)), "normalised_error_max": float(nerr.max()),
                              "note": "|dF_ij|/sqrt(F_ii F_jj); the reference's own FMA-vs-non-FMA noise floor is 2e-6..5e-6"}
            line["speedup_vs_cpu"] = line["value"] / line["cpu_baseline"]["value"]
    except Exception as exc:
        line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
