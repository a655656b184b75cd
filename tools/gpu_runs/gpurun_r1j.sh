#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 1 4 5; do python bench.py --steps 30 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['method'], 'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'roof %.3f'%d['roofline']['frac'],'kms %.4f'%d['roofline']['kernel_ms'], 'ms/step %.4f'%d['ms_per_step'])"; done
