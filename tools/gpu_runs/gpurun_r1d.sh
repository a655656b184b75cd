#!/bin/bash
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for c in 2 1 4 5; do python bench.py --steps 20 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_v2_cfg$c.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_v2_cfg$c.json').read()); print(d['config']['method'], 'value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'])"; done
ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o gpurun_out/prof_loglike_pv2_v2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
