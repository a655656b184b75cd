#!/bin/bash
# Fisher stencil with the carrier shared across detectors and across the RA/DEC/psi stencil points (slim variant):
# Fisher parity tests, Fisher bench, launch list
mkdir -p gpurun_out/r2k
export GWAT_B200_LIB=$PWD/variants/slim/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_sampler_gpu.py -m gpu -q -k "fisher or Fisher" 2>&1 | tail -8
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 > gpurun_out/r2k/bench_fisher.json 2>&1; tail -c 900 gpurun_out/r2k/bench_fisher.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2k/launches_fisher.csv python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 > gpurun_out/r2k/fisher_run.log 2>&1
