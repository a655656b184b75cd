#!/bin/bash
# round 2, eleventh call: cooperative setup, populate_source per role branch (coopA) or hoisted (coopB), against one thread per walker (base)
bash tools/gpu_runs/ab.sh r2_11 "base coopA coopB" "1 2 4"
