#!/bin/bash
# round 2, call 37 (1 GPU): the device sampler with the library as committed -- bench line and the launch list of its steps
O=gpurun_out/r2_37; mkdir -p $O
python bench.py --gpus 1 --workload sampler --steps 200 --warmup 20 > $O/bench_sampler_1gpu.json 2> $O/bench_sampler_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_sampler.csv python bench.py --gpus 1 --workload sampler --steps 30 --warmup 10 > $O/ncu_launch_run.log 2>&1
python - "$O" <<'PY'
import csv, json, sys
from collections import defaultdict
O = sys.argv[1]
d = json.loads(open(O + "/bench_sampler_1gpu.json").read().strip().splitlines()[-1])
print("sampler 1 GPU: %.4g %s, e2e %.4g, ms/step %.4f" % (d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"]), d.get("swap_exchange"))
rows = list(csv.reader(open(O + "/launches_sampler.csv")))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
kn, mv = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
t = defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) > mv:
        name = r[kn].split("<")[0].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        t[name][0] += 1
        t[name][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print("%-32s %5d launches %10.1f us  %5.1f %%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
PY
