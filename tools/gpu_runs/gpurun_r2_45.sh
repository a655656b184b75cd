#!/bin/bash
# round 2, call 45 (1 GPU): ncu --set full of k_setup (IMRPhenomPv2 with the repack dealt out over the roles; IMRPhenomD) and the launch lists
# of cfg1 / cfg2 with the last library
O=gpurun_out/r2_45; mkdir -p /tmp/prof $O
for c in 2 1; do
  ncu --set full --clock-control none --import-source on -k regex:k_setup -s 4 -c 1 -o /tmp/prof/setup_cfg$c -f python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline --no-extras > $O/ncu_full_setup_cfg$c.log 2>&1
  ncu -i /tmp/prof/setup_cfg$c.ncu-rep --page raw --csv > $O/setup_cfg${c}_raw.csv 2>/dev/null
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_cfg$c.csv python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
done
python - $O <<'PY'
import csv, sys
O = sys.argv[1]
for c in (2, 1):
    rows = list(csv.reader(open("%s/setup_cfg%d_raw.csv" % (O, c))))
    d = dict(zip(rows[0], rows[-1]))
    print("cfg%d k_setup: %s us, regs %s, issue %s %%, warps active %s" % (c, d["gpu__time_duration.sum"], d["launch__registers_per_thread"], d["smsp__issue_active.avg.pct_of_peak_sustained_active"], d["smsp__warps_active.avg.per_cycle_active"]))
PY
