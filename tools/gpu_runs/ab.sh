#!/bin/bash
# usage (on the GPU box): bash tools/gpu_runs/ab.sh <tag> "<variant> <variant> ..." ["<configs>"]  -- A/B of kernel-experiment libraries
# (variants/<name>/libgwat_b200.so, "shipped" = the in-tree library): k_loglike ms and ms/step per config, twice each.
tag=$1; variants=$2; configs=${3:-"1 2 4"}
O=gpurun_out/$tag
mkdir -p $O
for rep in 1 2; do
for v in $variants; do
  if [ "$v" = shipped ]; then unset GWAT_B200_LIB; else export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so; fi
  for c in $configs; do
    python bench.py --config $c --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/${v}_cfg${c}_$rep.json 2> $O/${v}_cfg${c}_$rep.err
    python - "$v" "$c" "$rep" "$O/${v}_cfg${c}_$rep.json" <<'PY'
import json, sys
v, c, rep, path = sys.argv[1:]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print("%-14s cfg%s rep%s  k_loglike %.4f ms  step %.4f ms  e2e %.4f ms  checksum %.17g" % (v, c, rep, d["roofline"]["kernel_ms"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["logL_checksum"]))
except Exception as e:
    print(v, c, rep, "FAILED", e)
PY
  done
done
done
