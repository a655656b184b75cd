#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
mkdir -p /tmp/prof
for c in 2 1; do
ncu --set full --clock-control none --import-source on -k regex:k_setup_mcmc -s 4 -c 1 -o /tmp/prof/setup_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline > gpurun_out/ncu_setup_cfg$c.log 2>&1
  ncu -i /tmp/prof/setup_cfg$c.ncu-rep --page raw --csv > gpurun_out/setup_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/setup_cfg$c.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/setup_cfg${c}_source.csv 2>/dev/null
done
gzip -f gpurun_out/*_source.csv
ls -la gpurun_out | tail -8
