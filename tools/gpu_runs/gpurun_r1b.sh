#!/bin/bash
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --config 1 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 2 -o gpurun_out/prof_loglike_pv2_r1b -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
