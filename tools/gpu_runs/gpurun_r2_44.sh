#!/bin/bash
# round 2, call 44 (N GPUs): the driver's torchrun line with the topology-aware core pinning (default for several ranks) and without it
N=${1:-2}
O=gpurun_out/r2_44_$N; mkdir -p $O
lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket" > $O/lscpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29644"
for rep in 1 2; do
for mode in pinned unpinned; do
  if [ $mode = unpinned ]; then X=--no-pin-cores; else X=; fi
  $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras $X > $O/bench_${mode}_$rep.json 2> $O/bench_${mode}_$rep.err
  python -c "
import json
d=json.loads([l for l in open('$O/bench_${mode}_$rep.json').read().strip().splitlines() if l.startswith('{')][-1])
print('$mode rep$rep n=%d value %.4g e2e %.4g' % (d['n_gpus'], d['value'], d['e2e']['value']), ['%.4f' % x for x in d['e2e']['per_rank_ms_per_step']], d['e2e'].get('host_cores_per_rank'))"
done
done
cat $O/lscpu.txt
