#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -12
for c in 2 1 4 5; do python bench.py --steps 30 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['method'], 'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'roof %.3f'%d['roofline']['frac'],'kms %.4f'%d['roofline']['kernel_ms'], 'ms/step %.4f'%d['ms_per_step'])"; done
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 2>&1 | tail -1 | cut -c1-260
python tools/bench_sampler.py --config 2 --lanes 2 --no-fisher 2>&1 | tail -1 | cut -c1-330
