#!/bin/bash
# 8 GPUs of one box: weak-scaling bench line (cfg2 per GPU), the driver's launch line
mkdir -p gpurun_out/r2e
nvidia-smi --query-gpu=index,name --format=csv | head -10
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/r2e/bench_8gpu.json 2> gpurun_out/r2e/bench_8gpu.err
tail -c 1500 gpurun_out/r2e/bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2e/bench_4gpu.json 2> gpurun_out/r2e/bench_4gpu.err
tail -c 600 gpurun_out/r2e/bench_4gpu.json
