#!/bin/bash
# round 2, fifteenth call: cosmologies, equatorial_orientation / horizon_coord, sky-averaged Fishers of the ppE / gIMR sets on the GPU
O=gpurun_out/r2_15; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest.log
tail -30 $O/pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > $O/bench_default.json 2> $O/bench_default.err; head -c 400 $O/bench_default.json; echo
