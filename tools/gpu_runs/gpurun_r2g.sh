#!/bin/bash
# sampler lanes with the bin kernel capped at 3 CTAs/SM (slim variant) vs the full library (4 CTAs/SM everywhere)
mkdir -p gpurun_out/r2g
for v in slim full; do
  if [ $v = slim ]; then export GWAT_B200_LIB=$PWD/variants/slim/libgwat_b200.so; else unset GWAT_B200_LIB; fi
  for c in 2 1 4; do
    python tools/bench_sampler.py --config $c --lanes 2 --no-fisher > gpurun_out/r2g/${v}_cfg${c}_gauss.json 2>&1
    python tools/bench_sampler.py --config $c --lanes 2 --deferred 1 --warmup 400 --steps 300 > gpurun_out/r2g/${v}_cfg${c}_full.json 2>&1
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2g/*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print("%-44s %.4g chain-steps/s  ms/step %.4f" % (f, d["value"], d["ms_per_step"]))
    except Exception as e:
        print(f, "ERR", open(f).read()[-300:])
PY
