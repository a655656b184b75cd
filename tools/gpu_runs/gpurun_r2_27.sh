#!/bin/bash
# round 2, call 27 (1 GPU): k_loglike at 5 and 6 CTAs per SM (48 / 40 registers) against the shipped 4 (64 registers)
bash tools/gpu_runs/ab.sh r2_27 "shipped c5 c6" "1 2 4"
