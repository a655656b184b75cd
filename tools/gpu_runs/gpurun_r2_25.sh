#!/bin/bash
# round 2, call 25 (1 GPU): GPU tier with the equatorial-orientation Fishers, the default bench line, and the ncu evidence
# refreshed against the library as committed (twist-up algebra of 6aff3b5 and later): launch list + --set full of k_loglike
O=gpurun_out/r2_25; mkdir -p /tmp/prof $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launch_run.log 2>&1
for c in 2 1 4 5; do
  ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o /tmp/prof/loglike_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline --no-extras > $O/ncu_full_cfg$c.log 2>&1
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page raw --csv > $O/loglike_cfg${c}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/loglike_cfg$c.ncu-rep --page details --csv > $O/loglike_cfg${c}_details.csv 2>/dev/null
done
ncu -i /tmp/prof/loglike_cfg2.ncu-rep --page source --csv --print-source sass,cuda > $O/loglike_cfg2_source.csv 2>/dev/null
gzip -f $O/*_source.csv
python -c "
import json
d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1])
print('default: value %.4g e2e %.4g ms/step %.4f kernel_ms %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline'].get('kernel_ms')))"
ls $O | wc -l
