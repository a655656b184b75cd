"""Dynamic temperature allocation (arXiv:1501.05823): the update rule against the reference's compiled
update_temperatures_full_ensemble, and the tuning loop on the device sampler."""
import numpy as np
import pytest

from gw_analysis_tools_b200 import sampler as smp, workloads


def _ladder(n_ens, n_t, tmax, rng):
    one = np.concatenate([[1.0], np.sort(rng.uniform(1.2, tmax * 0.9, n_t - 2)), [tmax]])
    return np.tile(one, n_ens)


@pytest.mark.parametrize("seed", range(6))
def test_update_rule_vs_reference(oracle, seed):
    from oracle import ref_sampler
    rng = np.random.default_rng(seed)
    n_ens, n_t = int(rng.integers(1, 5)), int(rng.integers(3, 9))
    temps = _ladder(n_ens, n_t, 40.0, rng)
    for t in (5, 250, 5000):
        A = rng.integers(0, 2, temps.size)
        A[0] = 0
        ref = ref_sampler.update_temperatures(temps, A, 1000, 10, t)
        got = smp.update_temperatures(temps, A, 1000, 10, t)
        assert np.array_equal(got, ref)
        # T = 1 chains and every ensemble's hottest chain stay, the ladder stays ordered inside an ensemble
        assert np.array_equal(got[temps == 1.0], temps[temps == 1.0]) and np.array_equal(got[temps == 40.0], temps[temps == 40.0])
        temps = got


@pytest.mark.gpu
def test_dynamic_allocation_on_the_device_sampler(ctx):
    wl = workloads.make(1, W=8, L=1024)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    n_ens, n_t = 4, 6
    temps0 = np.tile(np.geomspace(1.0, 200.0, n_t), n_ens)
    rng = np.random.default_rng(3)
    init = wl.inj[None, :] * (1 + 1e-4 * rng.normal(size=(temps0.size, wl.P)))
    prior = smp.prior_for(wl)
    s = smp.Sampler(ctx, wl.method, temps0, init, prior, wl.gmst, wl.T_segment, wl.mod, seed=5, swp_freq=5, fisher_exist=0, lanes=1)
    # the tuning loop of the library ...
    n = s.dynamic_temperatures(400, nu=10, t0=100)
    assert n == 79  # blocks while t < N_steps - swp_freq
    t1 = s.temperatures()
    # ... against the same loop driven from here with the pieces
    s2 = smp.Sampler(ctx, wl.method, temps0, init, prior, wl.gmst, wl.T_segment, wl.mod, seed=5, swp_freq=5, swap_rate=2.0, fisher_exist=0, lanes=1)
    temps, t = temps0.copy(), 0
    while t < 400 - 5:
        s2.run(5)
        t += 5
        A = np.concatenate([[0], s2.last_swap_accepts()])
        temps = smp.update_temperatures(temps, A, 100, 10, t)
        s2.set_temperatures(temps)
    assert np.array_equal(t1, temps)
    assert np.array_equal(s.state()[0], s2.state()[0])
    # the ends of every ensemble stayed, the interior moved, the ladder is still ordered
    lad = t1.reshape(n_ens, n_t)
    assert np.all(lad[:, 0] == 1.0) and np.all(lad[:, -1] == 200.0) and np.all(np.diff(lad, axis=1) > 0)
    assert not np.array_equal(t1, temps0)
    # cold chains may not be re-labelled
    bad = t1.copy()
    bad[0] = 1.5
    with pytest.raises(Exception):
        s.set_temperatures(bad)
    s.close()
    s2.close()
