#!/bin/bash
# round 2, call 39 (1 GPU): k_setup with the repack's libm calls dealt out over the roles (su1) against a slim build of the previous
# kernel (su0): ms/step of cfg1 / cfg2 / cfg4 / cfg5-short, with kernel timing of k_setup from a launch list
bash tools/gpu_runs/ab.sh r2_39 "su0 su1" "1 2 4"
for v in su0 su1; do
  export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so
  for c in 1 2; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_39/launches_${v}_cfg$c.csv python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
    python - gpurun_out/r2_39/launches_${v}_cfg$c.csv $v $c <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
kn, mv = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
v = [float(r[mv].replace(",", "")) for r in rows[h + 1:] if len(r) > mv and "k_setup" in r[kn]]
print("%s cfg%s k_setup under ncu: %.1f us (n=%d)" % (sys.argv[2], sys.argv[3], sum(v) / len(v) / 1e3, len(v)))
PY
  done
done
