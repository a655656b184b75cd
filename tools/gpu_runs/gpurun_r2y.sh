#!/bin/bash
# refresh of the ncu evidence against the library as committed: launch list (cfg2), --set full of k_loglike for the four configs
O=gpurun_out/r2y; mkdir -p /tmp/prof $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/ncu_launch_run.log 2>&1
for c in 2 1 4 5; do
  ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o /tmp/prof/loglike_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline > $O/ncu_full_cfg$c.log 2>&1
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page raw --csv > $O/loglike_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page details --csv > $O/loglike_cfg${c}_details.csv 2>/dev/null
done
ncu -i /tmp/prof/loglike_cfg2.ncu-rep --page source --csv --print-source sass,cuda > $O/loglike_cfg2_source.csv 2>/dev/null
gzip -f $O/*_source.csv
ls $O | wc -l
