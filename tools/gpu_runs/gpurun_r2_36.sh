#!/bin/bash
# round 2, call 36 (1 GPU): compute-sanitizer over small invocations of every kernel family (tools/sanitize_small.py)
O=gpurun_out/r2_36; mkdir -p $O
python tools/sanitize_small.py > $O/plain.log 2>&1; tail -3 $O/plain.log
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $O/$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small done|Invalid|hazard" $O/$tool.log | head -12
done
