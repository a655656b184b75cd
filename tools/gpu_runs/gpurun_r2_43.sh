#!/bin/bash
# round 2, call 43 (N GPUs of one box; N = first argument): weak and strong scaling of the likelihood path and of the sharded sampler
N=${1:-8}
O=gpurun_out/r2_43_$N; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643"
$TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_driver_line.json 2> $O/bench_driver_line.err
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras --pin-cores > $O/bench_weak_pinned.json 2> $O/bench_weak_pinned.err
$TR bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-extras --pin-cores > $O/bench_strong.json 2> $O/bench_strong.err
NCCL_DEBUG=INFO $TR bench.py --gpus $N --workload sampler --scaling strong --steps 200 --warmup 20 > $O/bench_sampler_strong.json 2> $O/bench_sampler_strong.err
$TR bench.py --gpus $N --workload sampler --scaling weak --steps 200 --warmup 20 > $O/bench_sampler_weak.json 2> $O/bench_sampler_weak.err
python - "$O" <<'PY'
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(os.path.basename(f), "n=%d value %.4g %s e2e %.4g ms/step %.4f scaling %s" % (d["n_gpus"], d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d["scaling"]),
              {k: d[k] for k in ("swap_exchange",) if k in d}, d["e2e"].get("per_rank_ms_per_step"))
    except Exception as e:
        print(os.path.basename(f), "FAILED", e)
PY
grep -h "NCCL INFO.*\(NVLS\|via P2P\|Connected all\)" $O/bench_sampler_strong.err | head -4
