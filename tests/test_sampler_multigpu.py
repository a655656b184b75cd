"""Two GPUs: the sharded sampler reproduces the single-GPU device sampler bit for bit -- both the exchange inside the library
(gwat_b200_sampler_attach_ranks: ncclAllGather + device sweep, no host round trip) and the Python-level one
(ensemble.DistributedSampler over torch.distributed).  Skipped on boxes with fewer than two GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["cabi", "python"])
def test_two_ranks_reproduce_one(tmp_path, mode):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from gw_analysis_tools_b200 import engine, workloads
    from gw_analysis_tools_b200 import sampler as smp
    steps = 47
    out = str(tmp_path / "two.npz")
    worker = os.path.join(ROOT, "tests", "mgpu_sampler_worker.py")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29533" if mode == "python" else "29534", worker, out, str(steps), mode], check=True, timeout=600)
    two = np.load(out)
    # the same run on one GPU, swaps done by the device sweep
    ctx = engine.Context(0)
    wl = workloads.make(2, W=64, L=2048)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    temps = np.tile(np.geomspace(1.0, 30.0, 8), 4)
    init = wl.inj[None, :] + 0.2 * (wl.params[:32] - wl.inj[None, :])
    s = smp.Sampler(ctx, wl.method, temps, init, smp.prior_for(wl), wl.gmst, wl.T_segment, wl.mod, seed=9, swp_freq=3, history_length=20,
                    fisher_update_number=7, lanes=2)
    s.run(steps)
    pos, ll, lp = s.state()
    ct, _ = s.counters()
    assert np.array_equal(two["pos"], pos) and np.array_equal(two["ll"], ll) and np.array_equal(two["lp"], lp)
    assert np.array_equal(two["swap_accept"], ct["swap_accept"]) and np.array_equal(two["swap_reject"], ct["swap_reject"])
    assert ct["swap_accept"].sum() > 0
    if mode == "cabi":
        assert int(two["sweeps"]) > 0 and float(two["swap_ms"]) > 0
    s.close()  # before its context
    ctx.close()
