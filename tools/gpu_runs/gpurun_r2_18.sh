#!/bin/bash
# round 2, call 18: fused Fisher kernel with the transposed, padded, double-buffered derivative tile (variant fz) against the committed slim build
O=gpurun_out/r2_18; mkdir -p $O
for rep in 1 2; do
for v in coopA fz; do
  export GWAT_B200_LIB=$PWD/variants/$v/libgwat_b200.so
  python bench.py --config 3 --steps 3 --warmup 1 --no-extras --no-cpu-baseline --fisher-sources 50000 > $O/${v}_cfg3_$rep.json 2> $O/${v}_cfg3_$rep.err
  python -c "
import json,sys
d=json.loads(open('$O/${v}_cfg3_$rep.json').read().strip().splitlines()[-1])
print('$v rep$rep value %.5g e2e %.5g ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
done
export GWAT_B200_LIB=$PWD/variants/fz/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "fisher or Fisher" 2>&1 | tail -4
python tools/fisher_noise_report.py --sources 64 > $O/fisher_noise_fz.json 2> $O/fisher_noise_fz.err; head -c 600 $O/fisher_noise_fz.json; echo
