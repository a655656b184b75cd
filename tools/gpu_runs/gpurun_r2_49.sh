#!/bin/bash
# round 2, call 49 (1 GPU): the round-end sequence again with the last library (gwat_b200_detector_site, link-level calculate_snr): smoke(), the GPU tier, both bench arms
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
O=gpurun_out/r2_49; mkdir -p $O
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --gpus 1 > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json
for n in ('bench_reference_arm','bench_default'):
    d=json.loads(open('$O/'+n+'.json').read().strip().splitlines()[-1])
    print(n, 'value %.5g e2e %.5g %s' % (d['value'], d['e2e']['value'], d['unit']), d.get('clocks'), d.get('gpu_launches'))"
