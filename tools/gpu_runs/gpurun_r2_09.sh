#!/bin/bash
# round 2, ninth call: cooperative setup kernel (warp = role) against the same slim build with one thread per walker
bash tools/gpu_runs/ab.sh r2_09 "base coop" "1 2 4 5"
export GWAT_B200_LIB=$PWD/variants/coop/libgwat_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_queue.py -m gpu -q -x 2>&1 | tail -5
for c in 1 2; do
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/r2_09/launches_coop_cfg$c.csv \
    python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_09/ncu_launch_cfg$c.log 2>&1
grep -h "k_setup" gpurun_out/r2_09/launches_coop_cfg$c.csv | head -2 | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-
done
