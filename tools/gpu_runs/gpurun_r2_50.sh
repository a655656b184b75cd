#!/bin/bash
# round 2, call 50 (1 GPU): one bench line per BASELINE config (1, 3, 4, 5; 2 is the default line of call 49) with the last committed library
O=gpurun_out/r2_50; mkdir -p $O
for c in 1 3 4 5; do
  timeout 300 python bench.py --gpus 1 --config $c > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err
done
python - $O <<'PY'
import json, sys
O = sys.argv[1]
for c in (1, 3, 4, 5):
    try:
        d = json.loads(open("%s/bench_cfg%d.json" % (O, c)).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print("cfg%d value %.5g e2e %.5g %s | ms/step %.4g | roofline frac %.3g pipe %s | cpu %.5g (%s cores) | clocks %s %s" % (
            c, d["value"], d["e2e"]["value"], d["unit"], d["ms_per_step"], r.get("frac", float("nan")), r.get("fp64_pipe_active_pct_ncu"),
            d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
    except Exception as e:
        print("cfg%d FAILED %r" % (c, e))
PY
