#!/bin/bash
# round 2, call 30 (1 GPU): k_fisher_fused with the products of tile t-1 formed by warps 0-2 while the others differentiate tile t (ff3) against ff2
bash tools/gpu_runs/fisher_ab.sh r2_30 "ff2 ff3"
