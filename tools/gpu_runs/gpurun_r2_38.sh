#!/bin/bash
# round 2, call 38 (1 GPU): k_setup with the repack's libm calls dealt out over the roles (su1) against the shipped kernel:
# ms/step of cfg1 / cfg2 / cfg4 (checksums must be identical), then the parity tests that go through the repack
bash tools/gpu_runs/ab.sh r2_38 "shipped su1" "1 2 4"
export GWAT_B200_LIB=$PWD/variants/su1/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_sampler_gpu.py -m gpu -q -x -k "mcmc or repack or loglike or trajector or batch" 2>&1 | tail -4
