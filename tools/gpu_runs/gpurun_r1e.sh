#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for lib in "" variants/libgwat_b200_minb3.so; do
  echo "=== lib=$lib"
  for c in 2 1 5; do GWAT_B200_LIB=${lib:+$PWD/$lib} python bench.py --steps 20 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['method'], 'value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'])"; done
done
