#!/bin/bash
# round 2, fourth call (2 GPUs): sharded sampler through the library's NCCL exchange vs one GPU (bit-identical), strong/weak scaling lines
O=gpurun_out/r2_04
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
python -m pytest tests/test_sampler_multigpu.py -m gpu -q 2>&1 | tail -30 > $O/pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > $O/bench_weak_2.json 2> $O/bench_weak_2.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --no-extras > $O/bench_strong_2.json 2> $O/bench_strong_2.err
NCCL_DEBUG=INFO $TR bench.py --gpus 2 --workload sampler --scaling strong --steps 200 --warmup 20 > $O/bench_sampler_strong_2.json 2> $O/bench_sampler_strong_2.err
$TR bench.py --gpus 2 --workload sampler --scaling weak --steps 200 --warmup 20 > $O/bench_sampler_weak_2.json 2> $O/bench_sampler_weak_2.err
python bench.py --gpus 1 --workload sampler --steps 200 --warmup 20 > $O/bench_sampler_1.json 2> $O/bench_sampler_1.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --walkers 2048 > $O/bench_half_1.json 2> $O/bench_half_1.err
tail -12 $O/pytest_mgpu.log
for f in bench_weak_2 bench_strong_2 bench_sampler_strong_2 bench_sampler_weak_2 bench_sampler_1 bench_half_1; do echo "== $f"; tail -1 $O/$f.json | head -c 900; echo; grep -v "NCCL INFO" $O/$f.err | tail -3; done
grep -c "NCCL INFO" $O/bench_sampler_strong_2.err; grep "NCCL INFO.*\(NVLS\|P2P\|via\)" $O/bench_sampler_strong_2.err | head -5
