#!/bin/bash
# Fisher derivative kernel looping over 4 tiles per CTA: Fisher tests, bench (2048 and 1e5 sources), whole GPU tier
O=gpurun_out/r2t; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 > $O/bench_fisher.json 2>/dev/null; tail -c 900 $O/bench_fisher.json | cut -c1-420; echo
python tools/bench_fisher.py --sources 100000 --bins 4096 --cpu-sample 64 > $O/bench_fisher_100k.json 2>/dev/null; tail -c 900 $O/bench_fisher_100k.json | cut -c1-420; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_fisher.csv python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 > $O/fisher_run.log 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg2_full_steady.json 2>&1
python tools/bench_sampler.py --config 4 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg4_full_steady.json 2>&1
tail -c 300 $O/sampler_cfg2_full_steady.json | head -c 200
