#!/bin/bash
# pinned stats word: whole GPU tier, bench lines for the four configs (final numbers of the session)
O=gpurun_out/r2w; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in 2 1 4 5; do python bench.py --config $c > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; python -c "
import json; d=json.loads(open('$O/bench_cfg$c.json').read().strip().split('\n')[-1]); print('cfg$c value %.5g e2e %.5g ms/step %.4f e2e_ms %.4f k_ms %.4f frac %.3f cpu %.5g' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['cpu_baseline']['value']), d['clocks'])"; done
