#!/bin/bash
mkdir -p /tmp/prof
ncu --set full --clock-control none --import-source on -k regex:k_setup_mcmc -s 4 -c 1 -o /tmp/prof/setup -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_setup_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o /tmp/prof/loglike -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
for k in setup loglike; do
  ncu -i /tmp/prof/$k.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$k.ncu-rep --page details --csv > gpurun_out/${k}_details.csv 2>/dev/null
  ncu -i /tmp/prof/$k.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/${k}_source.csv 2>/dev/null
done
gzip -f gpurun_out/*_source.csv
ls -la gpurun_out
