#!/bin/bash
# round 2, seventh call: L1 prefetch of the next bin's table lines in k_loglike (variant pf1) against the same slim build without it
bash tools/gpu_runs/ab.sh r2_07 "base pf1" "1 2 4 5"
python -m pytest tests/test_dropin_link.py -m gpu -q 2>&1 | tail -3
