#!/bin/bash
# round 2, fourteenth call: shipped library with the cooperative setup kernel, the finish fused into k_loglike, opt-in kernel events
O=gpurun_out/r2_14; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $O/pytest.log
tail -6 $O/pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_14/bench_default.json").read().strip().splitlines()[-1])
print("cfg2 value %.4g e2e %.4g ms/step %.4f e2e ms %.4f kernel_ms %.4f frac %.3f launches/step %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d.get("gpu_launches_per_step")))
for k, v in d.get("extra", {}).items():
    try:
        print(k, "value %.4g e2e %.4g ms/step %s" % (v["value"], v["e2e"]["value"], v.get("ms_per_step")))
    except Exception as e:
        print(k, str(v)[:200])
PY
tail -3 $O/bench_default.err
python bench.py --workload sampler --steps 200 --warmup 20 > $O/bench_sampler.json 2> $O/bench_sampler.err; head -c 700 $O/bench_sampler.json; echo; tail -3 $O/bench_sampler.err
