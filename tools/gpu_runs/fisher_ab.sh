#!/bin/bash
# usage (on the GPU box): bash tools/gpu_runs/fisher_ab.sh <tag> "<variant> ..."  -- A/B of k_fisher_fused builds on cfg3 (50000 sources), twice each,
# then the Fisher parity tests with each variant library ("shipped" = the in-tree one)
tag=$1; variants=$2
O=gpurun_out/$tag; mkdir -p $O
for rep in 1 2; do
for v in $variants; do
  if [ "$v" = shipped ]; then unset GWAT_B200_LIB; else export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so; fi
  python bench.py --config 3 --fisher-sources 50000 --steps 3 --warmup 1 --no-extras --no-cpu-baseline > $O/${v}_cfg3_$rep.json 2> $O/${v}_cfg3_$rep.err
  python - "$v" "$rep" "$O/${v}_cfg3_$rep.json" <<'PY'
import json, sys
v, rep, path = sys.argv[1:]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print("%-10s rep%s  %.1f Fisher/s  e2e %.1f  ms/step %.2f  kernel_ms %s" % (v, rep, d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("roofline", {}).get("kernel_ms")))
except Exception as e:
    print(v, rep, "FAILED", e)
PY
done
done
for v in $variants; do
  if [ "$v" = shipped ]; then continue; else export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so; fi
  echo "== parity tests with $v"
  python -m pytest tests/test_gpu_parity.py tests/test_orientation.py tests/test_intrinsic.py tests/test_gpu_fullsize.py -m gpu -q -k "fisher or Fisher" 2>&1 | tail -4
done
