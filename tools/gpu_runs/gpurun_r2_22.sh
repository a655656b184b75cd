#!/bin/bash
# round 2, call 22 (1 GPU): the swap sweep's sequential walk replaced by runs + pointer doubling -- sampler tests (against the compiled
# reference and the host sweep, adversarial ladders included), then the 1-GPU sampler line
python -m pytest tests/test_sampler_gpu.py tests/test_sampler_reference.py tests/test_dynamic_temperatures.py -m gpu -q 2>&1 | tail -5
mkdir -p gpurun_out/r2_22
python bench.py --gpus 1 --workload sampler --steps 200 --warmup 20 > gpurun_out/r2_22/bench_sampler_1gpu.json 2> gpurun_out/r2_22/bench_sampler_1gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_22/bench_sampler_1gpu.json').read().strip().splitlines()[-1])
print('sampler 1 GPU: %.4g chain-steps/s, ms/step %.4f, swap %s' % (d['value'], d['ms_per_step'], d['swap_exchange']))"
python - <<'PY'
# the sweep alone at ladder sizes of 1, 2 and 8 GPUs' worth of chains: device (parallel / forced sequential) timed by wall clock around the hook
import time, numpy as np
from gw_analysis_tools_b200 import sampler as smp
from gw_analysis_tools_b200.engine import Context
ctx = Context(0)
for n in (4096, 8192, 32768, 131072):
    temps = np.tile(np.geomspace(1.0, 50.0, 8), n // 8)
    ll = 1e4 + 3 * np.random.default_rng(n).standard_normal(n)
    for mode in (0, 1):
        smp.swap_sweep_device(ctx, ll, temps, 1, 1, mode)
        t0 = time.perf_counter()
        for k in range(20): smp.swap_sweep_device(ctx, ll, temps, 1, k, mode)
        print("sweep hook n=%d mode=%d: %.3f ms per call (incl. 7 cudaMalloc/Free + copies)" % (n, mode, (time.perf_counter() - t0) / 20 * 1e3))
PY
