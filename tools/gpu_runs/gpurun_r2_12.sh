#!/bin/bash
# round 2, twelfth call: finish fused into k_loglike (single-chunk grids), counter zeroed by k_setup, programmatic dependent launch
O=gpurun_out/r2_12; mkdir -p $O
bash tools/gpu_runs/ab.sh r2_12 "coopA pdl" "1 2 4"
export GWAT_B200_LIB=$PWD/variants/pdl/libgwat_b200.so
echo "== no PDL"; GWAT_B200_NO_PDL=1 python bench.py --config 1 --steps 30 --warmup 5 --no-extras --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['logL_checksum'])"
echo "== no events"; GWAT_B200_NO_KERNEL_EVENTS=1 python bench.py --config 1 --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>&1 | tail -3 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['logL_checksum'])"
GWAT_B200_NO_KERNEL_EVENTS=1 python bench.py --config 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>&1 | tail -3 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['logL_checksum'])"
