#!/bin/bash
# round 2, call 35 (1 GPU): what the driver runs at round end, with the library as committed -- smoke(), the GPU tier, both bench arms
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
O=gpurun_out/r2_35; mkdir -p $O
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --gpus 1 > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json
for n in ('bench_reference_arm','bench_default'):
    d=json.loads(open('$O/'+n+'.json').read().strip().splitlines()[-1])
    print(n, 'value %.5g e2e %.5g %s' % (d['value'], d['e2e']['value'], d['unit']), d.get('clocks'), d.get('gpu_launches'))"
