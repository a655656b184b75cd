#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in "" variants/libgwat_b200_t256_m2.so variants/libgwat_b200_t128_m5.so variants/libgwat_b200_t128_m6.so variants/libgwat_b200_t256_m4.so variants/libgwat_b200_t512_m1.so; do
  echo "=== lib=$lib"
  for c in 2 1 5; do GWAT_B200_LIB=${lib:+$PWD/$lib} python bench.py --steps 20 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['method'], 'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'roof %.3f'%d['roofline']['frac'],'kms %.4f'%d['roofline']['kernel_ms'], 'ms/step %.4f'%d['ms_per_step'])"; done
done
ncu --set full --clock-control none --import-source on -k regex:k_setup_mcmc -s 4 -c 1 -o gpurun_out/prof_setup_pv2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_setup_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_loglike -s 4 -c 1 -o gpurun_out/prof_loglike_pv2_v3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
