#!/bin/bash
# round 2, call 32 (1 GPU): full GPU tier with the Fisher kernel changes (sqrt(wq) table, point flags in bit masks), the default bench line
# (cfg3 in `extra` at 10^5 sources), bench --config 3 and ncu --set full of k_fisher_fused
O=gpurun_out/r2_32; mkdir -p /tmp/prof $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --config 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
ncu --set full --clock-control none --import-source on -k regex:k_fisher_fused -s 1 -c 1 -o /tmp/prof/fisher_fused -f python tools/bench_fisher.py --sources 5000 --bins 4096 --cpu-sample 0 > $O/ncu_fused.log 2>&1
ncu -i /tmp/prof/fisher_fused.ncu-rep --page raw --csv > $O/fisher_fused_raw.csv 2>/dev/null
ncu -i /tmp/prof/fisher_fused.ncu-rep --page details --csv > $O/fisher_fused_details.csv 2>/dev/null
ncu -i /tmp/prof/fisher_fused.ncu-rep --page source --csv --print-source sass,cuda > $O/fisher_fused_source.csv 2>/dev/null
gzip -f $O/fisher_fused_source.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_fisher.csv python tools/bench_fisher.py --sources 20000 --bins 4096 --cpu-sample 0 > $O/ncu_launch_run.log 2>&1
python -c "
import json
for n in ('bench_default','bench_cfg3'):
    d=json.loads(open('$O/'+n+'.json').read().strip().splitlines()[-1])
    print(n, 'value %.5g e2e %.5g %s ms/step %.4f' % (d['value'], d['e2e']['value'], d['unit'], d['ms_per_step']), {k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('extra',{}).items()})"
