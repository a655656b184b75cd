#!/bin/bash
# round 2, call 41 (1 GPU): autocorrelation lengths on the device (batched cuFFT) against the compiled reference, with a timing at the
# size of a production run (512 cold chains x 15 parameters x 20000 steps)
python -m pytest tests/test_autocorr.py tests/test_chain_io.py tests/test_abi.py -m gpu -q 2>&1 | tail -6
python - <<'PY'
import time, numpy as np
from gw_analysis_tools_b200 import engine, sampler
ctx = engine.Context(0)
rng = np.random.default_rng(0)
pos = np.cumsum(rng.standard_normal((512, 20000, 15)), axis=1) * 0.01 + rng.standard_normal((512, 20000, 15))
sampler.autocorrelation_lengths(ctx, pos[:8])
t0 = time.perf_counter(); ac, tau = sampler.autocorrelation_lengths(ctx, pos, begin=1000); dt = time.perf_counter() - t0
print("autocorrelation lengths of %d rows x %d steps: %.3f s (%.0f rows/s), tau median %.1f" % (ac.size, 19000, dt, ac.size / dt, np.median(tau)))
PY
