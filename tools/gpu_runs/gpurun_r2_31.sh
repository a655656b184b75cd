#!/bin/bash
# round 2, call 31 (1 GPU): k_fisher_fused (sqrt table + flag masks = ff4) at 80 registers x 2 CTAs against 128 registers x 1 CTA per SM
O=gpurun_out/r2_31; mkdir -p $O
export GWAT_B200_LIB=$PWD/variants/ff4/libgwat_b200.so
for rep in 1 2; do
for mode in normal wide; do
  if [ $mode = wide ]; then export GWAT_B200_FISHER_WIDE=1; else unset GWAT_B200_FISHER_WIDE; fi
  python bench.py --config 3 --fisher-sources 50000 --steps 3 --warmup 1 --no-extras --no-cpu-baseline > $O/${mode}_$rep.json 2> $O/${mode}_$rep.err
  python -c "
import json
d=json.loads(open('$O/${mode}_$rep.json').read().strip().splitlines()[-1])
print('$mode rep$rep %.1f Fisher/s e2e %.1f' % (d['value'], d['e2e']['value']))"
done
done
