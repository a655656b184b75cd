#!/bin/bash
# round 2, call 24 (8 GPUs): the sharded sampler again with the ranks entering the timed call together (clock sampler started before the
# barrier) -- strong and weak; swap exchange reported as max over ranks and for the last-arriving rank
N=${1:-8}
O=gpurun_out/r2_24_$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus $N --workload sampler --scaling strong --steps 200 --warmup 20 > $O/bench_sampler_strong.json 2> $O/bench_sampler_strong.err
$TR bench.py --gpus $N --workload sampler --scaling weak --steps 200 --warmup 20 > $O/bench_sampler_weak.json 2> $O/bench_sampler_weak.err
$TR bench.py --gpus $N --workload sampler --scaling weak --steps 1000 --warmup 20 > $O/bench_sampler_weak_1000.json 2> $O/bench_sampler_weak_1000.err
python - "$O" <<'PY'
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(os.path.basename(f), "n=%d value %.4g %s e2e %.4g ms/step %.4f scaling %s" % (d["n_gpus"], d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d["scaling"]),
              {k: v for k, v in d["swap_exchange"].items() if k != "what"}, d["per_rank"])
    except Exception as e:
        print(os.path.basename(f), "FAILED", e)
PY
