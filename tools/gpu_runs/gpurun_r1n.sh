#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -3
mkdir -p /tmp/prof gpurun_out/final
for c in 2 1 4 5; do python bench.py --config $c > gpurun_out/final/bench_cfg$c.json 2> gpurun_out/final/bench_cfg$c.err; tail -c 600 gpurun_out/final/bench_cfg$c.json; echo; done
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 > gpurun_out/final/bench_fisher.json 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --no-fisher --cpu-sample 512 > gpurun_out/final/sampler_cfg2_gauss.json 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --lookahead 4 --warmup 600 --steps 400 > gpurun_out/final/sampler_cfg2_fisher_steady.json 2>&1
python tools/bench_sampler.py --config 1 --lanes 2 --no-fisher --cpu-sample 512 > gpurun_out/final/sampler_cfg1_gauss.json 2>&1
python tools/bench_sampler.py --config 1 --lanes 2 --lookahead 4 --warmup 600 --steps 400 > gpurun_out/final/sampler_cfg1_fisher_steady.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_launch_run.log 2>&1
for c in 2 1 4 5; do
  ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o /tmp/prof/loglike_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline > gpurun_out/final/ncu_full_cfg$c.log 2>&1
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page raw --csv > gpurun_out/final/loglike_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page details --csv > gpurun_out/final/loglike_cfg${c}_details.csv 2>/dev/null
done
ncu -i /tmp/prof/loglike_cfg2.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/final/loglike_cfg2_source.csv 2>/dev/null
gzip -f gpurun_out/final/*_source.csv
ls -la gpurun_out/final
