#!/bin/bash
# round 2, call 28 (1 GPU): ncu --set full + source page of k_fisher_fused as shipped (derivative tile Z[parameter][k], two buffers)
O=gpurun_out/r2_28; mkdir -p /tmp/prof $O
ncu --set full --clock-control none --import-source on -k regex:k_fisher_fused -s 1 -c 1 -o /tmp/prof/fisher_fused -f python tools/bench_fisher.py --sources 5000 --bins 4096 --cpu-sample 0 > $O/ncu_fused.log 2>&1
ncu -i /tmp/prof/fisher_fused.ncu-rep --page raw --csv > $O/fisher_fused_raw.csv 2>/dev/null
ncu -i /tmp/prof/fisher_fused.ncu-rep --page details --csv > $O/fisher_fused_details.csv 2>/dev/null
ncu -i /tmp/prof/fisher_fused.ncu-rep --page source --csv --print-source sass,cuda > $O/fisher_fused_source.csv 2>/dev/null
gzip -f $O/fisher_fused_source.csv
tail -3 $O/ncu_fused.log | cut -c1-300
ls -la $O
