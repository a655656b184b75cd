"""Worker of tests/test_sampler_multigpu.py: one rank of a sharded PTMCMC run (launched under torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gw_analysis_tools_b200 import engine, ensemble, workloads  # noqa: E402
from gw_analysis_tools_b200 import sampler as smp  # noqa: E402


def main():
    out, steps = sys.argv[1], int(sys.argv[2])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    ctx = engine.Context(local_rank)
    wl = workloads.make(2, W=64, L=2048)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    temps = np.tile(np.geomspace(1.0, 30.0, 8), 4)
    init = wl.inj[None, :] + 0.2 * (wl.params[:32] - wl.inj[None, :])
    s = ensemble.DistributedSampler(ctx, wl.method, temps, init, smp.prior_for(wl), wl.gmst, wl.T_segment, wl.mod, seed=9, swp_freq=3,
                                    history_length=20, fisher_update_number=7, lanes=2)
    s.run(steps)
    pos, ll, lp = s.state()
    if rank == 0:
        np.savez(out, pos=pos, ll=ll, lp=lp, swap_accept=s.swap_accept, swap_reject=s.swap_reject)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
