#!/bin/bash
# round 2, call 29 (1 GPU): k_fisher_fused with the sqrt(wq) table + four product accumulators (ff1) and with the point flags in bit masks (ff2)
bash tools/gpu_runs/fisher_ab.sh r2_29 "shipped ff1 ff2"
