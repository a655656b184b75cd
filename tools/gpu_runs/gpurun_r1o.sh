#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/bench_fisher.py --sources 2048 --bins 4096 --cpu-sample 64 2>&1 | tail -1 | cut -c1-700
for args in "--config 2 --lanes 2 --deferred 1 --warmup 600 --steps 200" "--config 2 --lanes 2 --deferred 0 --warmup 600 --steps 200" "--config 2 --lanes 2 --deferred 1 --warmup 20 --steps 200" "--config 1 --lanes 2 --deferred 1 --warmup 600 --steps 200" "--config 4 --lanes 2 --deferred 1 --warmup 100 --steps 100"; do
  echo "== $args"; python tools/bench_sampler.py $args 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.4f  launches/step %.1f acc %.2f swap %.2f fisher_updates %d nan %d active %s kms %s'%(d['value'],d['ms_per_step'],d['launches_per_step'],d['accept_fraction'],d['swap_accept_fraction'],d['fisher_updates'],d['fisher_nan'],d['active_bin_fraction'],d['k_loglike_ms_full_ensemble']))"
done
