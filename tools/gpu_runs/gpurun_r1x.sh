#!/bin/bash
# kernel experiment: early exit of empty chunks (slim), 4 CTAs/SM variant (c4); sweep of steps per thread
set -x
mkdir -p gpurun_out/r1x
run() { python bench.py --config $1 --no-cpu-baseline --steps 50 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4g e2e %.4g ms/step %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"; }
export GWAT_B200_LIB=$PWD/variants/slim/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_queue.py -m gpu -x -q -k "mcmc_batch_vs_golden or loglike_mcmc_vs_oracle or full_size or queue or pool or single_caller" 2>&1 | tail -5
for s in 4 8 16 32 64; do
  for c in 1 2 4 5; do
    echo "slim steps=$s cfg=$c"; GWAT_B200_STEPS_PER_THREAD=$s run $c
  done
done 2>&1 | grep -v "^+" | tee gpurun_out/r1x/sweep_slim.txt
export GWAT_B200_LIB=$PWD/variants/c4/libgwat_b200.so
for s in 16 64; do
  for c in 1 2 4 5; do
    echo "c4 steps=$s cfg=$c"; GWAT_B200_STEPS_PER_THREAD=$s run $c
  done
done 2>&1 | grep -v "^+" | tee gpurun_out/r1x/sweep_c4.txt
