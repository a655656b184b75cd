#!/bin/bash
# kernel experiment: per-CTA seed tables + W-independent chunking (slim variant library), sweep of steps per thread
set -x
mkdir -p gpurun_out/r1w
export GWAT_B200_LIB=$PWD/variants/slim/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mcmc_batch_vs_golden or loglike_mcmc_vs_oracle or full_size" 2>&1 | tail -5
for s in 2 4 8 16 32 64; do
  for c in 1 2 4 5; do
    echo "steps=$s cfg=$c"
    GWAT_B200_STEPS_PER_THREAD=$s python bench.py --config $c --no-cpu-baseline --steps 50 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4g e2e %.4g ms/step %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
  done
done 2>&1 | tee gpurun_out/r1w/sweep.txt
