#!/bin/bash
# round 2, call 51 (1 GPU): ncu launch list of the default bench command with the last committed library, and the device sampler's bench line
O=gpurun_out/r2_51; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 200 python tools/bench_sampler.py > $O/bench_sampler.json 2> $O/bench_sampler.err
python - $O <<'PY'
import csv, sys, collections, json
O = sys.argv[1]
rows = [r for r in csv.reader(open(O + "/launches_default.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: x = float(r[v].replace(",", ""))
    except ValueError: continue
    n = r[k].split("(")[0][:60]; t[n][0] += 1; t[n][1] += x
tot = sum(b for a, b in t.values())
for n, (a, b) in sorted(t.items(), key=lambda kv: -kv[1][1])[:8]:
    print("%-60s launches %4d  total %10.1f us  share %5.1f %%" % (n, a, b / 1e3 if tot > 1e6 else b, 100 * b / tot))
try:
    print(open(O + "/bench_sampler.json").read().strip().splitlines()[-1][:600])
except Exception as e:
    print("sampler bench:", e, open(O + "/bench_sampler.err").read()[-400:])
PY
