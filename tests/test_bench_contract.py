"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`) runs the
reference's own CPU code (oracle/_ref) and prints one line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(oracle):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().split("\n")[-1])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["impl"] == "reference"
    assert base["metric"].startswith(line["metric"].split(" (")[0])
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "evals/s" and line["higher_is_better"] is True and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["warmup"] >= 3 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert "workload" in line["config"] and "model" not in line["config"]


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0  # no CPU fallback: the product arm fails loudly
    assert "reference" not in out.stdout


def test_core_groups_partition_the_affinity_mask():
    """bench.py deals the host's physical cores out over the ranks of a multi-GPU run: the groups (hyperthread siblings together)
    must partition the CPUs this process may run on, and eight ranks must get disjoint, non-empty shares when there are enough cores."""
    import importlib
    import os
    import sys
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    bench = importlib.import_module("bench")
    cpus = set(os.sched_getaffinity(0))
    groups = bench.physical_core_groups(cpus)
    flat = [c for g in groups for c in g]
    assert sorted(flat) == sorted(cpus) and len(flat) == len(set(flat))
    world = min(8, len(groups))
    share = len(groups) // world
    shares = [set(c for g in groups[r * share:(r + 1) * share] for c in g) for r in range(world)]
    assert all(shares) and sum(len(s) for s in shares) == len(set().union(*shares))
