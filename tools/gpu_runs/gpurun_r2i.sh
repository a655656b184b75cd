#!/bin/bash
# kernel experiment: cp.async staging of the per-bin table values (variant stg) vs the current kernel (variant slim)
mkdir -p gpurun_out/r2i
run() { python bench.py --config $1 --no-cpu-baseline --steps 50 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4g e2e %.4g ms/step %.4f e2e_ms %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))"; }
export GWAT_B200_LIB=$PWD/variants/stg/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_queue.py tests/test_sampler_gpu.py -m gpu -x -q -k "mcmc_batch_vs_golden or loglike_mcmc_vs_oracle or full_size or queue or pool or single_caller or cfg1 or cfg2 or glq or gaussleg or odd_length" 2>&1 | tail -8
for v in stg slim stg slim; do
  export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so
  for c in 2 4 1 5; do echo "$v cfg=$c"; run $c; done
done 2>&1 | tee gpurun_out/r2i/bench.txt
