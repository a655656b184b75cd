#!/usr/bin/env python3
"""Where does the Fisher noise come from?  (VERDICT r1 weak #2.)  Run on the GPU box:  python tools/fisher_noise_report.py

For the cfg3 bench population (workloads.fisher_sources) it evaluates the order-4, 3-detector Fisher matrices four ways
  ref      the reference's code (oracle/_ref)
  self     the reference again on inputs moved by parts in 1e14, x4, and its FMA-contracted build when present
           (tests/fisher_noise.py: the reference's own reproducibility)
  harness  the kernels' mathematics compiled for the host (tests/_build/libgwat_host_harness.so: glibc libm, no FMA
           contraction) -- the same algorithm as the GPU, different libm and contraction
  gpu      the CUDA path through the C ABI
and prints one JSON line with the per-source max normalised differences |dF_ij|/sqrt(F_ii F_jj):
  gpu-ref, harness-ref, gpu-harness, self.  gpu-harness isolates CUDA's libm + FMA contraction; harness-ref the algebraic
  reformulation; self is the yardstick.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fisher_noise as FN  # noqa: E402
from gw_analysis_tools_b200 import engine, workloads  # noqa: E402
from oracle import gwat_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=64)
    ap.add_argument("--bins", type=int, default=4096)
    args = ap.parse_args()
    S = args.sources
    srcs = workloads.fisher_sources(S)
    f = 20.0 + 0.25 * np.arange(args.bins)
    dets = ["Hanford", "Livingston", "Virgo"]
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    nt = max(1, len(os.sched_getaffinity(0)))
    ref = gwat_ref.fisher_numerical_batch("IMRPhenomD", srcs, dets, f, psd, 11, order=4, nthreads=nt)
    self_diff = FN.reference_self_difference(gwat_ref, "IMRPhenomD", srcs, dets, f, psd, 11, 4, nthreads=nt)
    ctx = engine.Context(0)
    ctx.set_network(dets, f, psd)
    gpu = ctx.fisher_numerical_batch("IMRPhenomD", srcs, 11, order=4)
    hh = C.CDLL(os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so"))
    dp = C.POINTER(C.c_double)
    har = np.zeros_like(ref)
    for i, s in enumerate(srcs):
        for det in dets:
            out = np.zeros((11, 11))
            rc = hh.hh_fisher_numerical(b"IMRPhenomD", det.encode(), b"Hanford", 11, 4, C.byref(s), f.ctypes.data_as(dp), f.size,
                                        psd[0].ctypes.data_as(dp), out.ctypes.data_as(dp))
            assert rc == 0
            har[i] += out
    ok = np.all(np.isfinite(ref.reshape(S, -1)), axis=1) & np.all(np.isfinite(gpu.reshape(S, -1)), axis=1)

    def worst(a, b):
        return FN.normalised_error(a[ok], b[ok]).reshape(int(ok.sum()), -1).max(axis=1)
    g_r, h_r, g_h, sd = worst(gpu, ref), worst(har, ref), worst(gpu, har), self_diff[ok]

    def stats(x):
        return {"median": float(np.median(x)), "p90": float(np.quantile(x, 0.9)), "max": float(x.max())}
    line = {"sources": int(ok.sum()), "bins": args.bins,
            "max_normalised_difference_per_source": {"gpu_vs_reference": stats(g_r), "harness_vs_reference": stats(h_r),
                                                     "gpu_vs_harness": stats(g_h), "reference_vs_itself": stats(sd)},
            "ratio_to_reference_self_difference": {"gpu": stats(g_r / sd), "harness": stats(h_r / sd)},
            "median_entry_error": {"gpu": float(np.median(FN.normalised_error(gpu[ok], ref[ok]))),
                                   "harness": float(np.median(FN.normalised_error(har[ok], ref[ok])))}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
