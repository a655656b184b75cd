#!/bin/bash
# usage: tools/gpu_runs/submit.sh <name> <timeout_s> [--gpus N]   -- runs tools/gpu_runs/gpurun_<name>.sh on a GPU box, retrying while the pod is busy
name=$1; tmo=$2; shift 2
for try in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout $tmo "$@" -- "bash tools/gpu_runs/gpurun_${name}.sh" > gpurun_out/${name}.log 2>&1
  if grep -q "status=transient" gpurun_out/${name}.log; then sleep 90; continue; fi
  break
done
tail -80 gpurun_out/${name}.log
