#!/bin/bash
# final library with the restructured Fisher stencil: whole GPU tier, smoke, Fisher bench + launch list, samplers with Fisher, cfg2 bench
O=gpurun_out/r2m; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 > $O/bench_fisher.json 2>&1; tail -c 300 $O/bench_fisher.json; echo
python tools/bench_fisher.py --sources 20000 --bins 4096 --cpu-sample 0 > $O/bench_fisher_20k.json 2>&1; tail -c 400 $O/bench_fisher_20k.json | cut -c1-300; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_fisher.csv python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 > $O/fisher_run.log 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg2_full_steady.json 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --deferred 0 --warmup 600 --steps 400 > $O/sampler_cfg2_full_refsched.json 2>&1
python tools/bench_sampler.py --config 1 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg1_full_steady.json 2>&1
python tools/bench_sampler.py --config 4 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg4_full_steady.json 2>&1
python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -c 200 $O/bench_cfg2.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m/sampler_*.json")):
    d = json.loads(open(f).read().strip().split("\n")[-1]); print("%-50s %.4g chain-steps/s ms/step %.4f" % (f, d["value"], d["ms_per_step"]))
PY
