#!/bin/bash
# re-entry sanity run: GPU parity tier, smoke, default bench and the reference arm on the restored tree
set -x
mkdir -p gpurun_out/r1v
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r1v/bench_ref.json
python bench.py 2>&1 | tail -1 | tee gpurun_out/r1v/bench_cfg2.json
python bench.py --config 1 2>&1 | tail -1 | tee gpurun_out/r1v/bench_cfg1.json
