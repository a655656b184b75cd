#!/bin/bash
python -m pytest tests/test_sampler_gpu.py -m gpu -q 2>&1 | tail -5
for args in "--config 2 --lanes 1 --no-fisher" "--config 2 --lanes 2 --no-fisher" "--config 2 --lanes 1" "--config 2 --lanes 2" "--config 2 --lanes 2 --lookahead 4" "--config 1 --lanes 1 --no-fisher" "--config 1 --lanes 2 --no-fisher"  "--config 1 --lanes 2 --lookahead 4" "--config 4 --lanes 2 --steps 60 --lookahead 4"; do
  echo "== $args"; python tools/bench_sampler.py $args 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.4f  frac_of_like %.3f launches/step %.1f acc %.2f swap %.2f fisher_updates %d nan %d'%(d['value'],d['ms_per_step'],d['fraction_of_likelihood_ceiling'],d['launches_per_step'],d['accept_fraction'],d['swap_accept_fraction'],d['fisher_updates'],d['fisher_nan']))"
done
