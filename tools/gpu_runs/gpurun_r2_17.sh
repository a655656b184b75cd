#!/bin/bash
# round 2, call 17: the library as committed (cooperative setup, fused finish, orientation / cosmology / sky-averaged widening, dynamic
# temperatures): GPU tier, smoke, bench lines of all configs and of the reference arm, launch lists, ncu --set full of k_loglike
# (cfg1, cfg2, cfg4, cfg5), k_setup (cfg1, cfg2) and k_fisher_fused (cfg3)
O=gpurun_out/r2_17; mkdir -p /tmp/prof $O
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_ref.err; tail -c 300 $O/bench_reference_arm.json; echo
python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; head -c 300 $O/bench_default.json; echo
for c in 1 3 4 5; do python bench.py --config $c --no-extras > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; head -c 250 $O/bench_cfg$c.json; echo; done
python bench.py --masses light --no-extras > $O/bench_cfg2_light.json 2> $O/bench_cfg2_light.err; head -c 250 $O/bench_cfg2_light.json; echo
for c in 1 2; do
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 16 --csv --log-file $O/launches_cfg$c.csv \
    python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launch_cfg$c.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file $O/launches_fisher.csv \
    python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline --no-extras --fisher-sources 20000 > $O/ncu_launch_fisher.log 2>&1
for c in 2 1 4 5; do
  ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o /tmp/prof/loglike_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline --no-extras > $O/ncu_full_cfg$c.log 2>&1
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page raw --csv > $O/loglike_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page details --csv > $O/loglike_cfg${c}_details.csv 2>/dev/null
done
ncu -i /tmp/prof/loglike_cfg2.ncu-rep --page source --csv --print-source sass,cuda > $O/loglike_cfg2_source.csv 2>/dev/null
for c in 2 1; do
  ncu --set full --clock-control none --import-source on -k regex:k_setup -s 4 -c 1 -o /tmp/prof/setup_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline --no-extras > $O/ncu_full_setup_cfg$c.log 2>&1
  ncu -i /tmp/prof/setup_cfg$c.ncu-rep --page raw --csv > $O/setup_cfg${c}_raw.csv 2>/dev/null
done
ncu -i /tmp/prof/setup_cfg2.ncu-rep --page source --csv --print-source sass,cuda > $O/setup_cfg2_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_fisher_fused -s 1 -c 1 -o /tmp/prof/fisher -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --fisher-sources 20000 > $O/ncu_full_fisher.log 2>&1
ncu -i /tmp/prof/fisher.ncu-rep --page raw --csv > $O/fisher_fused_raw.csv 2>/dev/null
ncu -i /tmp/prof/fisher.ncu-rep --page details --csv > $O/fisher_fused_details.csv 2>/dev/null
gzip -f $O/*_source.csv
ls $O | wc -l; du -sh $O
