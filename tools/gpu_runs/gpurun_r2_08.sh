#!/bin/bash
# round 2, eighth call: k_loglike loop variants (computed cutoff test, software-pipelined carrier inputs, 3 CTAs/SM)
bash tools/gpu_runs/ab.sh r2_08 "base v1 v2 v2m3" "1 2 4"
