#!/bin/bash
# ncu --set full of the Fisher derivative and assembly kernels (cfg3 shape), and a sky-averaged Fisher throughput line
O=gpurun_out/r2s; mkdir -p $O /tmp/prof
ncu --set full --clock-control none --import-source on -k regex:k_fisher_deriv -s 0 -c 1 -o /tmp/prof/fisher_deriv -f python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 > $O/ncu_deriv.log 2>&1
ncu -i /tmp/prof/fisher_deriv.ncu-rep --page raw --csv > $O/fisher_deriv_raw.csv 2>/dev/null
ncu -i /tmp/prof/fisher_deriv.ncu-rep --page details --csv > $O/fisher_deriv_details.csv 2>/dev/null
ncu --set full --clock-control none -k regex:k_fisher_assemble -s 0 -c 1 -o /tmp/prof/fisher_asm -f python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 > $O/ncu_asm.log 2>&1
ncu -i /tmp/prof/fisher_asm.ncu-rep --page raw --csv > $O/fisher_assemble_raw.csv 2>/dev/null
python - <<'PY'
import time, json, numpy as np, sys
sys.path.insert(0, ".")
from gw_analysis_tools_b200 import engine, abi
L = 3000
f = 15 + np.arange(L) * ((1000 - 15.) / (L - 1))
psd = engine.populate_noise(f, "Hanford_O1_fitted") ** 2
rng = np.random.default_rng(7)
S = 20000
srcs = []
for _ in range(S):
    m = np.sort(rng.uniform(3, 100, 2))[::-1]
    srcs.append(abi.source_defaults(mass1=m[0], mass2=m[1], Luminosity_Distance=rng.uniform(10, 1000), spin1=[0, 0, rng.uniform(-.9, .9)],
                                    spin2=[0, 0, rng.uniform(-.9, .9)], tc=rng.uniform(0, 10), phiRef=rng.uniform(0, 6.28), f_ref=20.0, sky_average=1))
arr = (abi.Source * S)(*srcs)
ctx = engine.Context(0)
ctx.set_network(["Hanford"], f, psd[None, :])
ctx.fisher_numerical_batch("IMRPhenomD", arr, 7, order=4, detector_index=0)
t0 = time.perf_counter(); F = ctx.fisher_numerical_batch("IMRPhenomD", arr, 7, order=4, detector_index=0); dt = time.perf_counter() - t0
print(json.dumps({"metric": "sky-averaged order-4 Fisher matrices/sec (IMRPhenomD, dim 7, 3000 bins, the set-up of the reference's testing/fisher_comparison.cpp)",
                  "value": S / dt, "unit": "Fisher/s", "sources": S, "seconds": dt, "device_ms": ctx.last_kernel_ms, "finite_fraction": float(np.mean(np.all(np.isfinite(F.reshape(S, -1)), axis=1)))}))
PY
