#!/bin/bash
# slim variant: step multipliers in shared memory (fewer spills) + snr_batch; parity subset, new SNR tests, bench
set -x
mkdir -p gpurun_out/r2c
run() { python bench.py --config $1 --no-cpu-baseline --steps 50 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4g e2e %.4g ms/step %.4f e2e_ms %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))"; }
export GWAT_B200_LIB=$PWD/variants/slim/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_queue.py tests/test_sampler_gpu.py tests/test_noise_snr.py -m gpu -x -q -k "mcmc_batch_vs_golden or loglike_mcmc_vs_oracle or full_size or queue or pool or single_caller or cfg1 or cfg2 or repack or antenna" 2>&1 | tail -8
for c in 1 2 4 5; do echo "slim cfg=$c"; run $c; done 2>&1 | grep -v "^+" | tee gpurun_out/r2c/bench.txt
