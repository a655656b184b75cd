#!/bin/bash
# BASELINE config 3 at its full size: 10^5 sources; which sources give non-finite matrices, and does the reference agree?
mkdir -p gpurun_out/r2n
python tools/bench_fisher.py --sources 100000 --bins 4096 --cpu-sample 64 > gpurun_out/r2n/bench_fisher_100k.json 2> gpurun_out/r2n/err.log; tail -c 1200 gpurun_out/r2n/bench_fisher_100k.json; tail -3 gpurun_out/r2n/err.log
