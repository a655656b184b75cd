#!/bin/bash
# kernel v7b, product library as committed: full GPU test tier, smoke, bench lines (all configs, reference arm), Fisher and
# sampler benches, launch lists, ncu --set full of k_loglike for every BASELINE config and of k_setup_mcmc (cfg1, cfg2)
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
O=gpurun_out/r2h; mkdir -p /tmp/prof $O
python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_arm_cfg2.json 2> $O/bench_ref.err; tail -c 300 $O/bench_reference_arm_cfg2.json; echo
for c in 2 1 4 5; do python bench.py --config $c > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; tail -c 200 $O/bench_cfg$c.json; echo; done
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 > $O/bench_fisher.json 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --no-fisher --cpu-sample 512 > $O/sampler_cfg2_gauss.json 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg2_full_steady.json 2>&1
python tools/bench_sampler.py --config 2 --lanes 2 --deferred 0 --warmup 600 --steps 400 > $O/sampler_cfg2_full_refsched.json 2>&1
python tools/bench_sampler.py --config 1 --lanes 2 --no-fisher --cpu-sample 512 > $O/sampler_cfg1_gauss.json 2>&1
python tools/bench_sampler.py --config 1 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg1_full_steady.json 2>&1
python tools/bench_sampler.py --config 4 --lanes 2 --no-fisher > $O/sampler_cfg4_gauss.json 2>&1
python tools/bench_sampler.py --config 4 --lanes 2 --deferred 1 --warmup 600 --steps 400 > $O/sampler_cfg4_full_steady.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/ncu_launch_run.log 2>&1
for c in 2 1 4 5; do
  ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o /tmp/prof/loglike_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline > $O/ncu_full_cfg$c.log 2>&1
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page raw --csv > $O/loglike_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page details --csv > $O/loglike_cfg${c}_details.csv 2>/dev/null
done
ncu -i /tmp/prof/loglike_cfg2.ncu-rep --page source --csv --print-source sass,cuda > $O/loglike_cfg2_source.csv 2>/dev/null
for c in 2 1; do
  ncu --set full --clock-control none --import-source on -k regex:k_setup_mcmc -s 4 -c 1 -o /tmp/prof/setup_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline > $O/ncu_full_setup_cfg$c.log 2>&1
  ncu -i /tmp/prof/setup_cfg$c.ncu-rep --page raw --csv > $O/setup_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/setup_cfg$c.ncu-rep --page source --csv --print-source sass,cuda > $O/setup_cfg${c}_source.csv 2>/dev/null
done
gzip -f $O/*_source.csv
ls $O | wc -l
