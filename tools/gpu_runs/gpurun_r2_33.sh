#!/bin/bash
# round 2, call 33 (1 GPU): k_fisher_fused with 2 x 2 blocked inner products (ff5) against the shipped kernel (one pair per thread)
bash tools/gpu_runs/fisher_ab.sh r2_33 "shipped ff5"
