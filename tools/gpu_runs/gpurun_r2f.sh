#!/bin/bash
# bench with the NVML clock sampler: cfg2 (default) and cfg1; sampler benches with kernel v7b
mkdir -p gpurun_out/r2f
python -m pytest tests/test_noise_snr.py -m gpu -q 2>&1 | tail -3
python bench.py --no-cpu-baseline | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['clocks'])"
python bench.py --config 1 --no-cpu-baseline | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['clocks'])"
python tools/bench_sampler.py --config 2 --lanes 2 --no-fisher --cpu-sample 512 > gpurun_out/r2f/sampler_cfg2_gauss.json 2>&1; tail -c 700 gpurun_out/r2f/sampler_cfg2_gauss.json; echo
python tools/bench_sampler.py --config 2 --lanes 2 --deferred 1 --warmup 600 --steps 400 > gpurun_out/r2f/sampler_cfg2_full_steady.json 2>&1; tail -c 500 gpurun_out/r2f/sampler_cfg2_full_steady.json; echo
python tools/bench_sampler.py --config 1 --lanes 2 --no-fisher --cpu-sample 512 > gpurun_out/r2f/sampler_cfg1_gauss.json 2>&1; tail -c 500 gpurun_out/r2f/sampler_cfg1_gauss.json; echo
python tools/bench_sampler.py --config 1 --lanes 2 --deferred 1 --warmup 600 --steps 400 > gpurun_out/r2f/sampler_cfg1_full_steady.json 2>&1; tail -c 500 gpurun_out/r2f/sampler_cfg1_full_steady.json; echo
python tools/bench_sampler.py --config 2 --lanes 2 --deferred 0 --warmup 600 --steps 400 > gpurun_out/r2f/sampler_cfg2_full_refsched.json 2>&1; tail -c 500 gpurun_out/r2f/sampler_cfg2_full_refsched.json; echo
python tools/bench_sampler.py --config 4 --lanes 2 --deferred 1 --warmup 300 --steps 200 > gpurun_out/r2f/sampler_cfg4_full_steady.json 2>&1; tail -c 500 gpurun_out/r2f/sampler_cfg4_full_steady.json; echo
