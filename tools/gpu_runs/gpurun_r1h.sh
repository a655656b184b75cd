#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 1; do python bench.py --steps 50 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['method'], 'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'roof %.3f'%d['roofline']['frac'],'kms %.4f'%d['roofline']['kernel_ms'], 'ms/step %.4f'%d['ms_per_step'], d['clocks'])"; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 2>&1 | tail -2 | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference 2>&1 | tail -1 | cut -c1-300
