#!/bin/bash
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in 2 1 4 5; do python bench.py --steps 20 --warmup 3 --config $c 2>&1 | tail -1 > gpurun_out/bench_cfg$c.json; cat gpurun_out/bench_cfg$c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['method'], 'value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'],'cpu',d.get('cpu_baseline',{}).get('value'))"; done
