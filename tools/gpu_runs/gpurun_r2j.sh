#!/bin/bash
# Fisher path: launch list of tools/bench_fisher.py (which kernel dominates?), and the new drop-in / LOSC tests on the final library
mkdir -p gpurun_out/r2j
python -m pytest tests/test_losc.py tests/test_gwatpy_dropin.py tests/test_noise_snr.py -m gpu -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2j/launches_fisher.csv python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 > gpurun_out/r2j/fisher_run.log 2>&1
tail -c 400 gpurun_out/r2j/fisher_run.log
