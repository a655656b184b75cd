"""Kernel experiment helper: wall-clock ms per host-buffer likelihood call (gwat_b200_loglike_mcmc_batch) for a BASELINE config."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gw_analysis_tools_b200 import engine, workloads  # noqa: E402


def main():
    for cfg in [int(a) for a in sys.argv[1:]] or [1, 2]:
        wl = workloads.make(cfg)
        ctx = engine.Context(0)
        ctx.set_network(wl.detectors, wl.f, wl.psd, np.zeros((wl.D, wl.L), dtype=complex))
        for _ in range(20):
            out = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter()
            for _ in range(200):
                out = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
            best = min(best, (time.perf_counter() - t0) / 200 * 1e3)
        print("cfg%d  %.4f ms per call (best of 5 x 200)  checksum %.17g" % (cfg, best, float(np.nansum(out))))
        ctx.close()


if __name__ == "__main__":
    main()
