#!/bin/bash
python -m pytest tests/test_maximized.py -m gpu -q 2>&1 | tail -25
python -m pytest tests -m gpu -q --deselect tests/test_maximized.py 2>&1 | tail -3
for c in 2 1 4 5; do python bench.py --steps 30 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['method'], 'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'roof %.3f'%d['roofline']['frac'],'kms %.4f'%d['roofline']['kernel_ms'], 'ms/step %.4f'%d['ms_per_step'])"; done
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 0 2>&1 | tail -1 | cut -c1-300
mkdir -p /tmp/prof
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_loglike -s 4 -c 1 --csv python bench.py --steps 3 --warmup 3 --config 5 --no-cpu-baseline 2>/dev/null | tail -5
