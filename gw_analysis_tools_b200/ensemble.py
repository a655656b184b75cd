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


class DistributedSampler:
    """The device-resident PTMCMC sampler (``sampler.Sampler``) with its chains sharded contiguously over the ranks of a
    ``torch.distributed`` group -- one process per GPU.  Steps need no communication.  At every swap interval the ranks
    all-gather logL (8 B per chain) and, when the sweep happens, positions and priors (8 (P+1) B per chain), run the
    reference's sequential sweep over the WHOLE ladder on the host (``gwat_b200_swap_sweep_host``: the single-GPU sweep's
    draws and arithmetic) and keep their own slots.  Random draws are functions of the global chain index, so N ranks
    reproduce the single-GPU run bit for bit (tests/test_sampler_multigpu.py)."""

    DRAW_SWAP_GATE = 5

    def __init__(self, ctx, method, temps, initial_positions, prior, gmst, T_segment, mod=None, group=None, **options):
        import torch.distributed as dist
        from . import sampler as smp
        self._smp = smp
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        temps = np.asarray(temps, dtype=np.float64)
        init = np.asarray(initial_positions, dtype=np.float64)
        self.C, self.P = init.shape
        if self.C % self.world != 0:
            raise ValueError("DistributedSampler needs the number of chains to be a multiple of the world size")
        n = self.C // self.world
        self.lo, self.hi = self.rank * n, (self.rank + 1) * n
        self.temps = temps
        self.swp_freq = int(options.get("swp_freq", 5))
        self.swap_rate = float(options.pop("swap_rate", 1.0 / self.swp_freq))
        self.seed = int(options.get("seed", 1))
        self.local = smp.Sampler(ctx, method, temps[self.lo:self.hi], init[self.lo:self.hi], prior, gmst, T_segment, mod,
                                 swap_rate=0.0, chain_index_offset=self.lo, **options)
        self.sweep = 0
        self.since = 0
        self.swap_accept = np.zeros(self.C, dtype=np.int64)
        self.swap_reject = np.zeros(self.C, dtype=np.int64)

    def _all_gather(self, local):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return local
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.as_tensor(np.ascontiguousarray(local), device=dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return torch.cat(out).cpu().numpy()

    def _sweep(self):
        gate, _ = self._smp.draw_uniform2(self.seed, self.sweep, 0, self.DRAW_SWAP_GATE)
        if gate < self.swap_rate and self.C > 1:
            pos, ll, lp = self.local.state()
            all_ll = self._all_gather(ll)
            src, acc = self._smp.swap_sweep_host(all_ll, self.temps, self.seed, self.sweep)
            if acc.any():
                all_pos, all_lp = self._all_gather(pos), self._all_gather(lp)
                mine = src[self.lo:self.hi]
                self.local.set_state(all_pos[mine], all_ll[mine], all_lp[mine])
            for which, sel in ((self.swap_accept, acc == 1), (self.swap_reject, acc == 0)):
                which[:-1] += sel
                which[1:] += sel
        self.sweep += 1

    def run(self, n_steps):
        while n_steps > 0:
            k = min(n_steps, self.swp_freq - self.since)
            self.local.run(k)
            self.since += k
            n_steps -= k
            if self.since == self.swp_freq:
                self._sweep()
                self.since = 0

    def state(self):
        """Global (positions, logL, logP), gathered."""
        pos, ll, lp = self.local.state()
        return self._all_gather(pos), self._all_gather(ll), self._all_gather(lp)
