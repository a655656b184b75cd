#!/bin/bash
# end-of-session check of the library as committed: whole GPU tier, smoke, default bench line, reference arm, Fisher bench
O=gpurun_out/r2u; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 3 2>/dev/null | tail -1 | cut -c1-200
python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; python -c "
import json; d=json.loads(open('$O/bench_cfg2.json').read().strip().split('\n')[-1]); print('cfg2', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])"
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 2>/dev/null | tail -1 | cut -c1-260
