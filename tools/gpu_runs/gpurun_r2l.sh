#!/bin/bash
# Fisher deriv with cached finishes: 3 CTAs/SM (variant slim) vs 2 CTAs/SM (variant f2)
mkdir -p gpurun_out/r2l
for v in slim f2; do
  export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so
  python -m pytest tests/test_gpu_parity.py tests/test_sampler_gpu.py -m gpu -q -k "fisher or Fisher" 2>&1 | tail -2
  python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 > gpurun_out/r2l/bench_fisher_$v.json 2>&1; tail -c 700 gpurun_out/r2l/bench_fisher_$v.json | cut -c1-420; echo
  python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 2>&1 | tail -1 | cut -c1-260
done
