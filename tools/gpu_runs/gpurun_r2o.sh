#!/bin/bash
# kernel experiment: Estrin split of the atan polynomial (variant est) vs current (variant slim)
mkdir -p gpurun_out/r2o
run() { python bench.py --config $1 --no-cpu-baseline --steps 50 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4g e2e %.4g ms/step %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"; }
export GWAT_B200_LIB=$PWD/variants/est/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mcmc_batch_vs_golden or loglike_mcmc_vs_oracle or full_size or waveform_vs_oracle" 2>&1 | tail -3
for v in est slim est slim; do
  export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so
  for c in 2 4 1; do echo "$v cfg=$c"; run $c; done
done 2>&1 | tee gpurun_out/r2o/bench.txt
