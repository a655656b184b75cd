#!/bin/bash
# first GPU contact: parity tests + a crude timing of cfg1/cfg2-shaped PhenomD batches
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np
from gw_analysis_tools_b200 import engine, workloads
from oracle import gwat_ref as R
c = engine.Context(0)
for cfg, W, L in [(1, 1024, 8192), (1, 4096, 16384)]:
    for masses in [(36., 29.), (10., 8.)]:
        wl = workloads.make(cfg, W=W, L=L, masses=masses)
        _, src = R.loglike_mcmc_batch(wl.method, wl.mod, wl.inj[None, :], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, None, return_sources=True)
        wl.data = R.coherent_response(wl.method, src[0], wl.detectors, wl.f)
        c.set_network(wl.detectors, wl.f, wl.psd, wl.data)
        for it in range(3):
            t = time.time(); ll = c.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment); dt = time.time() - t
        print(cfg, W, L, masses, "e2e ms", dt * 1e3, "kernel ms", c.last_kernel_ms, "active", c.last_active_bins, "evals/s", W / dt, "bins/s(kernel)", c.last_active_bins / (c.last_kernel_ms * 1e-3))
        t = time.time(); ref = R.loglike_mcmc_batch(wl.method, wl.mod, wl.params[:64], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data); dt = time.time() - t
        print("   cpu oracle", R.max_threads(), "threads: evals/s", 64 / dt, "max rel err", np.abs(ll[:64] - ref).max() / np.abs(ref).max())
PY
