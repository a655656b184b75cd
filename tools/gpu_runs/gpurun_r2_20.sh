#!/bin/bash
# round 2, call 20 (2 GPUs): the swap scan staged through shared memory -- sampler tests (single GPU against the compiled reference, two
# ranks bit-identical to one), then the scaling lines of call 19 at N = 2 and the 1-GPU sampler line
python -m pytest tests/test_sampler_multigpu.py tests/test_sampler_gpu.py tests/test_sampler_reference.py tests/test_dynamic_temperatures.py -m gpu -q 2>&1 | tail -5
bash tools/gpu_runs/gpurun_r2_19.sh 2
python bench.py --gpus 1 --workload sampler --steps 200 --warmup 20 > gpurun_out/r2_19_2/bench_sampler_1gpu.json 2> gpurun_out/r2_19_2/bench_sampler_1gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_19_2/bench_sampler_1gpu.json').read().strip().splitlines()[-1])
print('sampler 1 GPU: %.4g chain-steps/s, ms/step %.4f, swap %s' % (d['value'], d['ms_per_step'], d['swap_exchange']))"
