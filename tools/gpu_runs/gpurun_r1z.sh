#!/bin/bash
# kernel experiment: unit size 16 / 32 / 64 bins per thread (variants slim, u32, u64), default units per CTA
set -x
mkdir -p gpurun_out/r1z
run() { python bench.py --config $1 --no-cpu-baseline --steps 50 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4g e2e %.4g ms/step %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"; }
for v in slim u32 u64 slim u32 u64; do
  export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so
  for c in 1 2 4 5; do
    echo "$v cfg=$c"; run $c
  done
done 2>&1 | grep -v "^+" | tee gpurun_out/r1z/sweep.txt
