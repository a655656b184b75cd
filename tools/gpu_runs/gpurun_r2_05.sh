#!/bin/bash
# round 2, fifth call: link-level drop-in (libgwat_b200_dropin.so) and the new gwatpy symbols on the GPU; parallel swap sweep timing
O=gpurun_out/r2_05
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
python bench.py --workload sampler --steps 200 --warmup 20 > $O/bench_sampler.json 2> $O/bench_sampler.err
./tests/_build/dropin_caller_b200 > $O/dropin_caller_b200.txt 2> $O/dropin_caller_b200.err
tail -25 $O/pytest.log; head -c 600 $O/bench_default.json; echo; tail -3 $O/bench_default.err; head -c 1800 $O/bench_sampler.json; echo; tail -5 $O/bench_sampler.err; tail -5 $O/dropin_caller_b200.txt; tail -3 $O/dropin_caller_b200.err
