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
    mode = sys.argv[3] if len(sys.argv) > 3 else "python"   # "cabi": the exchange inside the library (gwat_b200_sampler_attach_ranks)
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
    if mode == "cabi":
        n = len(temps) // world
        lo = rank * n
        s = smp.Sampler(ctx, wl.method, temps[lo:lo + n], init[lo:lo + n], smp.prior_for(wl), wl.gmst, wl.T_segment, wl.mod, seed=9, swp_freq=3,
                        history_length=20, fisher_update_number=7, lanes=2, chain_index_offset=lo)
        if world > 1:
            ids = [smp.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            s.attach_ranks(ids[0], rank, world)
        s.run(steps // 2)
        s.run(steps - steps // 2)
        pos, ll, lp = s.state()
        ct, _ = s.counters()
        parts = [pos, ll, lp, ct["swap_accept"].astype(np.float64), ct["swap_reject"].astype(np.float64)]
        if world > 1:
            gathered = []
            for a in parts:
                t = torch.as_tensor(np.ascontiguousarray(a), device="cuda")
                o = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(o, t)
                gathered.append(torch.cat(o).cpu().numpy())
            parts = gathered
        if rank == 0:
            np.savez(out, pos=parts[0], ll=parts[1], lp=parts[2], swap_accept=parts[3].astype(np.int64), swap_reject=parts[4].astype(np.int64),
                     swap_ms=s.last_swap_ms, sweeps=s.last_sweeps)
        s.close()
    else:
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
