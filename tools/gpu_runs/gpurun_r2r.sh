#!/bin/bash
# final library (sky-averaged Fishers, amplitude/phase, plain-C API): whole GPU tier, smoke, default bench, Fisher bench
O=gpurun_out/r2r; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -c 250 $O/bench_cfg2.json; echo
python bench.py --config 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg1', d['value'], d['e2e']['value'], d['clocks'])"
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 2>/dev/null | tail -1 | cut -c1-330
