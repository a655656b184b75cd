#!/bin/bash
# full library (kernel v7b + snr_batch): whole GPU test tier; cfg5 repeatability, slim variant vs full library
mkdir -p gpurun_out/r2d
python -m pytest tests -m gpu -q 2>&1 | tail -8
run() { python bench.py --config $1 --no-cpu-baseline --steps ${2:-50} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4g e2e %.4g ms/step %.4f e2e_ms %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))"; }
for rep in 1 2; do
for c in 5 2 4 1; do echo "full cfg=$c"; run $c; done
export GWAT_B200_LIB=$PWD/variants/slim/libgwat_b200.so
for c in 5; do echo "slim cfg=$c"; run $c; done
unset GWAT_B200_LIB
done 2>&1 | tee gpurun_out/r2d/bench.txt
