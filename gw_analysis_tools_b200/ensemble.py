"""Multi-GPU plumbing for walker ensembles: sharding and the parallel-tempering swap exchange.

The hot path shards trivially -- every walker's likelihood is independent given the (replicated) grid, PSDs and data -- so
each rank owns a contiguous block of whole temperature rungs and NO collective runs on the likelihood path.  The one
exchange step of a parallel-tempered sampler is the swap between adjacent rungs: every rank needs the log-likelihoods
(and, for accepted swaps across a rank boundary, the positions) of its neighbour rungs.  That is an all-gather of
8*W bytes of logL plus 8*P*W bytes of positions -- latency-bound, done once per swap interval.
`torch.distributed` carries it (NCCL over NVLink on the GPUs, gloo in the CPU tests); it is plumbing, not compute.

Swap rule: the reference's single_chain_swap (src/mcmc_sampler_internals.cpp:1121-1184):
    accept iff  exp((ll1 - ll2)/T2 - (ll1 - ll2)/T1) >= u,   u ~ U(0,1)
"""
import numpy as np


def shard_rungs(n_temps, walkers_per_temp, rank, world):
    """Contiguous block of whole temperature rungs owned by `rank`: returns (first_rung, n_rungs, first_walker, n_walkers)."""
    base, extra = divmod(n_temps, world)
    n = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n, first * walkers_per_temp, n * walkers_per_temp


def swap_decisions(logl, temps, walkers_per_temp, uniforms, parity):
    """Decide swaps between rungs (t, t+1) for t = parity, parity+2, ... and walker slot k of each rung.

    logl[W] global log-likelihoods (rung-major), temps[n_temps], uniforms[n_pairs, walkers_per_temp] in [0,1).
    Returns a permutation `src` of the global walker indices: new_position[i] = old_position[src[i]].
    """
    n_temps = len(temps)
    W = n_temps * walkers_per_temp
    src = np.arange(W)
    pair = 0
    for t in range(parity, n_temps - 1, 2):
        T1, T2 = temps[t], temps[t + 1]
        a = np.arange(t * walkers_per_temp, (t + 1) * walkers_per_temp)
        b = a + walkers_per_temp
        if T1 != T2:
            d = logl[a] - logl[b]
            power = d / T2 - d / T1
            accept = np.exp(np.minimum(power, 700.0)) >= uniforms[pair]
            accept &= np.isfinite(power)
            src[a[accept]], src[b[accept]] = b[accept], a[accept]
        pair += 1
    return src


def exchange_and_swap(local_positions, local_logl, temps, walkers_per_temp, rank, world, step, seed=0, group=None):
    """One PT swap step over the whole ensemble.

    local_positions: torch tensor [W_local, P]; local_logl: torch tensor [W_local] (any device the process group supports).
    All ranks all-gather logL and positions, take the SAME swap decisions (shared counter-based RNG: seed, step) and keep
    their own rungs of the permuted ensemble.  Returns (new_local_positions, new_local_logl, n_accepted_global).
    Requires equal W_local on every rank (whole rungs, n_temps divisible by world).
    """
    import torch
    import torch.distributed as dist

    n_temps = len(temps)
    if n_temps % world != 0:
        raise ValueError("exchange_and_swap needs the number of temperature rungs to be a multiple of the world size")
    if world > 1:
        gl = [torch.empty_like(local_logl) for _ in range(world)]
        gp = [torch.empty_like(local_positions) for _ in range(world)]
        dist.all_gather(gl, local_logl.contiguous(), group=group)
        dist.all_gather(gp, local_positions.contiguous(), group=group)
        all_logl = torch.cat(gl)
        all_pos = torch.cat(gp)
    else:
        all_logl, all_pos = local_logl, local_positions
    rng = np.random.default_rng([seed, step])
    n_pairs = (n_temps - 1 - (step % 2) + 1) // 2
    uniforms = rng.random((max(n_pairs, 1), walkers_per_temp))
    src = swap_decisions(all_logl.detach().cpu().numpy(), np.asarray(temps, dtype=np.float64), walkers_per_temp, uniforms,
                         step % 2)
    _, _, w0, nw = shard_rungs(n_temps, walkers_per_temp, rank, world)
    idx = torch.as_tensor(src[w0:w0 + nw], device=all_pos.device)
    accepted = int((src != np.arange(src.size)).sum() // 2)
    return all_pos.index_select(0, idx), all_logl.index_select(0, idx), accepted
