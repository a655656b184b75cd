#!/bin/bash
# round 2, call 34 (2 GPUs): the tests that need two GPUs (sharded sampler, bit-identical to one GPU) and the driver's own torchrun line, with
# the library as committed at the end of the round
O=gpurun_out/r2_34; mkdir -p $O
python -m pytest tests/test_sampler_multigpu.py tests/test_multirank.py -m gpu -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu_driver_line.json 2> $O/bench_2gpu_driver_line.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_2gpu_reference_arm.json 2> $O/bench_2gpu_reference_arm.err
python - "$O" <<'PY'
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(os.path.basename(f), "n=%d value %.4g %s e2e %.4g ms/step %.4f" % (d["n_gpus"], d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"]), d["e2e"].get("per_rank_ms_per_step"), d.get("impl"))
    except Exception as e:
        print(os.path.basename(f), "FAILED", e)
PY
