#!/bin/bash
# round 2, call 21: twist-up algebra of the fused likelihood (Wigner sums as polynomials in cos/sin beta, one reciprocal root,
# slim power table above the inspiral, constant G for the IMRPhenomD families): GPU tier with the new library, then A/B against the
# library of commit c275377 (variants/base)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/gpu_runs/ab.sh r2_21 "base shipped" "1 2 4"
for v in base shipped; do
  if [ "$v" = shipped ]; then unset GWAT_B200_LIB; else export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so; fi
  python bench.py --config 2 --masses light --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_21/${v}_cfg2_light.json 2> gpurun_out/r2_21/${v}_cfg2_light.err
  python bench.py --config 5 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_21/${v}_cfg5.json 2> gpurun_out/r2_21/${v}_cfg5.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
for n in ("cfg2_light", "cfg5"):
    try:
        d = json.loads(open("gpurun_out/r2_21/%s_%s.json" % (v, n)).read().strip().splitlines()[-1])
        print("%-8s %-10s k_loglike %.4f ms  step %.4f ms  e2e %.4f ms  checksum %.17g" % (v, n, d["roofline"]["kernel_ms"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["logL_checksum"]))
    except Exception as e:
        print(v, n, "FAILED", e)
PY
done
