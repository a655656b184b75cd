#!/bin/bash
# kernel v7 as built into the product library: full GPU test tier, bench lines, launch list, ncu --set full of k_loglike (cfg2, cfg1) and k_setup_mcmc
python -m pytest tests -m gpu -q 2>&1 | tail -8
mkdir -p /tmp/prof gpurun_out/r2b
for c in 2 1 4 5; do python bench.py --config $c > gpurun_out/r2b/bench_cfg$c.json 2> gpurun_out/r2b/bench_cfg$c.err; tail -c 400 gpurun_out/r2b/bench_cfg$c.json; echo; done
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 > gpurun_out/r2b/bench_fisher.json 2>&1; tail -c 600 gpurun_out/r2b/bench_fisher.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/ncu_launch_run.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b/launches_cfg1.csv python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/ncu_launch_run1.log 2>&1
for c in 2 1; do
  ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o /tmp/prof/loglike_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline > gpurun_out/r2b/ncu_full_cfg$c.log 2>&1
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page raw --csv > gpurun_out/r2b/loglike_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page details --csv > gpurun_out/r2b/loglike_cfg${c}_details.csv 2>/dev/null
done
ncu -i /tmp/prof/loglike_cfg2.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r2b/loglike_cfg2_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_setup_mcmc -s 4 -c 1 -o /tmp/prof/setup_cfg2 -f python bench.py --steps 3 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r2b/ncu_full_setup_cfg2.log 2>&1
ncu -i /tmp/prof/setup_cfg2.ncu-rep --page raw --csv > gpurun_out/r2b/setup_cfg2_raw.csv 2>/dev/null
ncu -i /tmp/prof/setup_cfg2.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r2b/setup_cfg2_source.csv 2>/dev/null
gzip -f gpurun_out/r2b/*_source.csv
ls gpurun_out/r2b | wc -l
