"""GPU tier for the device-resident PTMCMC step, through the C ABI of include/gwat_b200_sampler.h.

* priors and Fisher eigen-systems against the restatement / the compiled reference,
* whole trajectories against the restated reference algorithm (oracle/ptmcmc_ref.py) driven by the SAME counter-based
  random numbers and by the compiled reference likelihood (oracle/_ref): positions must agree to 1e-9 step after step,
* invariants at scale: lane split does not change a single bit, posterior of an injected signal is recovered.
"""
import math

import numpy as np
import pytest

from gw_analysis_tools_b200 import sampler as smp
from gw_analysis_tools_b200 import workloads
from oracle import ptmcmc_ref as ref

pytestmark = pytest.mark.gpu


def _inject(ctx, wl):
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    wl.data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)
    return wl


def _ladder(n_ens, n_temps, tmax=30.0):
    return np.tile(np.geomspace(1.0, tmax, n_temps), n_ens)


def _start(wl, C, seed=3, spread=0.2):
    """Initial positions: the workload's scatter pulled towards the injection (finite prior, decent likelihood)."""
    p = wl.inj[None, :] + spread * (wl.params[:C] - wl.inj[None, :])
    return np.ascontiguousarray(p)


@pytest.mark.parametrize("cfg", [1, 2, 4, 5])
def test_log_prior_batch_vs_restatement(ctx, cfg):
    wl = workloads.make(cfg, W=512, L=1024 if cfg != 5 else 4096)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    prior = smp.prior_for(wl)
    rng = np.random.default_rng(cfg)
    params = wl.params.copy()
    # push a third of the walkers across some boundary
    k = rng.integers(0, wl.P, 170)
    params[np.arange(170), k] += rng.choice([-1, 1], 170) * 10.0
    got = smp.log_prior_batch(ctx, wl.method, params, prior, wl.mod)
    pd = prior.as_dict()
    want = np.array([ref.standard_log_prior(list(r), pd, "Pv2" in wl.method, "NRT" in wl.method) for r in params])
    both_inf = np.isneginf(got) & np.isneginf(want)
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    assert 50 < both_inf.sum() < 400
    fin = ~both_inf
    assert np.allclose(got[fin], want[fin], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("cfg", [1, 2])
def test_mcmc_fisher_vs_oracle(ctx, oracle, cfg):
    """MCMC_fisher_wrapper: sum over detectors of fisher_numerical("MCMC_"+method) + MCMC_fisher_transformations."""
    wl = workloads.make(cfg, W=4, L=4096)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    params = _start(wl, 4)
    F, vals, vecs = smp.mcmc_fisher_batch(ctx, wl.method, params, wl.gmst, order=4, mod=wl.mod)
    _, srcs = oracle.loglike_mcmc_batch(wl.method, wl.mod, params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, None, return_sources=True)
    for i in range(4):
        srcs[i].tc = wl.T_segment - srcs[i].tc  # undo the likelihood's flip: the Fisher is taken at the sampled tc
    total = oracle.fisher_numerical_batch("MCMC_" + wl.method, srcs, wl.detectors, wl.f, wl.psd, wl.P, 4, detector_index=-1, reference_index=0)
    for i in range(4):
        want = ref.fisher_transformations(total[i], cfg == 2)
        dg = np.sqrt(np.abs(np.diag(want)))
        nerr = np.abs(F[i] - want) / np.outer(dg, dg)
        assert np.median(nerr) <= 1e-6 and nerr.max() <= 1e-4, (i, np.median(nerr), nerr.max())
        # eigen-system of the matrix the device produced
        w = np.linalg.eigvalsh(F[i])
        assert np.allclose(vals[i], w, rtol=1e-7, atol=1e-9 * np.abs(w).max())
        assert np.allclose(vecs[i] @ vecs[i].T, np.eye(wl.P), atol=1e-11)
        assert np.allclose(vecs[i] @ F[i] @ vecs[i].T, np.diag(vals[i]), atol=1e-8 * np.abs(w).max())


def _restated(oracle, wl, temps, init, prior, seed, fisher_fn, **kw):
    pd = prior.as_dict()
    pv2, nrt = "Pv2" in wl.method, "NRT" in wl.method

    def ll(p):
        return oracle.loglike_mcmc_batch(wl.method, wl.mod, np.atleast_2d(p), wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data)
    return ref.Sampler(ll, lambda p: ref.standard_log_prior(list(p), pd, pv2, nrt), temps, init, seed, fisher=fisher_fn, **kw)


@pytest.mark.parametrize("cfg,fisher,lanes", [(1, False, 1), (1, True, 2), (2, True, 2), (4, True, 1), (5, True, 2)])
def test_trajectories_match_restated_reference(ctx, oracle, cfg, fisher, lanes):
    """Step for step against the restated reference algorithm running on the compiled reference likelihood.

    Fisher jumps: eigenvectors of a nearly degenerate, numerically differentiated matrix are not reproducible between two
    evaluations (1e-16 in the position comes back as 1e-6 in a vector), so the restatement jumps along the eigen-system the
    device holds for that chain -- its refresh schedule, the use of it, and everything else stay independent -- and the
    eigenvalues are checked against a stand-alone evaluation at the restatement's own position.  The matrices themselves
    are pinned to the compiled reference in test_mcmc_fisher_vs_oracle."""
    wl = _inject(ctx, workloads.make(cfg, W=64, L=1024 if cfg != 5 else 8192))
    C = 12
    temps = _ladder(3, 4, 20.0)
    init = _start(wl, C)
    prior = smp.prior_for(wl)
    kw = dict(swp_freq=3, history_length=12, history_update=2, fisher_update_number=4, check_stepsize_freq=5)
    seed = 77 + cfg
    g = smp.Sampler(ctx, wl.method, temps, init, prior, wl.gmst, wl.T_segment, wl.mod, seed=seed, fisher_exist=int(fisher), lanes=lanes, **kw)
    checked = []

    def fish(c, p):
        vals, vecs = g.fisher_state()
        _, v2, _ = smp.mcmc_fisher_batch(ctx, wl.method, p[None, :], wl.gmst, order=4, mod=wl.mod)
        big = np.abs(v2[0]) > 1e-6 * np.abs(v2[0]).max()
        assert np.allclose(vals[c][big], v2[0][big], rtol=1e-3), (c, vals[c], v2[0])
        checked.append(c)
        return vals[c], vecs[c]
    r = _restated(oracle, wl, temps, init, prior, seed, fish if fisher else None, **kw)
    pos, ll, lp = g.state()
    assert np.allclose(ll, r.ll, rtol=1e-9) and np.allclose(lp, r.lp, rtol=1e-13)
    n_steps = 42
    for step in range(n_steps):
        g.run(1)
        r.run(1)
        pos, ll, lp = g.state()
        assert np.allclose(pos, np.array(r.pos), rtol=1e-9, atol=1e-12), "diverged at step %d" % step
        assert np.allclose(ll, r.ll, rtol=1e-9)
        assert np.allclose(lp, r.lp, rtol=1e-9, atol=1e-9)
    ct, widths = g.counters()
    assert np.array_equal(ct["step_accept"], [c["step"][0] for c in r.ct]) and np.array_equal(ct["step_reject"], [c["step"][1] for c in r.ct])
    assert np.array_equal(ct["swap_accept"], [c["swap"][0] for c in r.ct]) and np.array_equal(ct["swap_reject"], [c["swap"][1] for c in r.ct])
    assert np.array_equal(ct["gauss_accept"] + ct["gauss_reject"], [sum(c["gauss"]) for c in r.ct])
    assert np.array_equal(ct["de_accept"] + ct["de_reject"], [sum(c["de"]) for c in r.ct])
    assert np.array_equal(ct["fisher_accept"] + ct["fisher_reject"], [sum(c["fisher"]) for c in r.ct])
    assert np.array_equal(ct["fisher_updates"] + ct["fisher_nan"], [c["fisher_updates"] for c in r.ct])
    assert np.allclose(widths, np.array(r.widths), rtol=1e-14)  # tuning is driven by integer counters only
    assert (ct["step_accept"] + ct["step_reject"] == n_steps).all()
    if fisher:
        assert ct["de_accept"].sum() + ct["de_reject"].sum() > 0 and ct["fisher_updates"].sum() > C and len(checked) > C
    assert ct["swap_accept"].sum() > 0


@pytest.mark.parametrize("cfg,fisher,lanes", [(1, False, 1), (1, True, 2), (2, True, 2), (4, True, 1), (5, True, 2)])
def test_trajectories_match_reference_steps(ctx, oracle, cfg, fisher, lanes):
    """Step for step against the reference's OWN compiled sampler (oracle/_ref: src/mcmc_sampler_internals.cpp,
    src/mcmc_sampler.cpp, src/standardPriorLibrary.cpp, unmodified), run by its own loop PTMCMC_MH_step_incremental on its own
    likelihood and priors, and fed -- through the scripted gsl_rng stand-in -- the counter-based draws the device uses, in the
    order the reference's code asks for them (oracle/ref_sampler.py).  Positions after every step (and after every swap sweep),
    likelihoods, priors, acceptance/swap counters and tuned widths must agree.

    Fisher jumps: the eigenvectors of a numerically differentiated, nearly degenerate matrix are not reproducible between two
    evaluations, so the reference's Eigen stand-in is told to return, for every refresh, the eigen-system the device holds after
    that step (its refresh schedule, its use of the system and everything else are the reference's own); the matrices themselves
    are pinned to the compiled reference in test_mcmc_fisher_vs_oracle."""
    from oracle import ref_sampler as rs
    if not hasattr(oracle.lib(), "oracle_sampler_create"):
        pytest.skip("oracle/_ref was built without the sampler translation units")
    wl = _inject(ctx, workloads.make(cfg, W=64, L=1024 if cfg != 5 else 8192))
    C = 12
    temps = _ladder(3, 4, 20.0)
    init = _start(wl, C)
    prior = smp.prior_for(wl)
    kw = dict(swp_freq=3, history_length=12, history_update=2, fisher_update_number=4, check_stepsize_freq=5)
    seed = 177 + cfg
    n_rounds = 14
    n_steps = n_rounds * kw["swp_freq"]
    g = smp.Sampler(ctx, wl.method, temps, init, prior, wl.gmst, wl.T_segment, wl.mod, seed=seed, fisher_exist=int(fisher), lanes=lanes, **kw)
    fisher_hist = {-1: g.fisher_state()}
    pos0, ll0, lp0 = g.state()
    traj, lls, lps = [], [], []
    for step in range(n_steps):
        g.run(1)
        p, l, q = g.state()
        traj.append(p)
        lls.append(l)
        lps.append(q)
        if fisher:
            fisher_hist[step] = g.fisher_state()
    R = rs.RefSampler(wl, temps, init, prior, seed, n_rounds, smp.draw_uniform2, ref.normal_from, fisher_exist=fisher,
                      initial_fisher=(lambda c: (fisher_hist[-1][0][c], fisher_hist[-1][1][c])) if fisher else None, **kw)
    kinds, refreshes = R.script(lambda s, c: (fisher_hist[s][0][c], fisher_hist[s][1][c]))
    R.run()
    res = R.results()
    R.close()
    d = res["diag"]
    assert d["rng_underflow"] == 0 and d["uniforms_left"] == 0 and d["normals_left"] == 0 and d["fisher_script_underflow"] == 0, d
    out = res["output"]
    assert np.allclose(res["ll"][:, 0], ll0, rtol=1e-9) and np.allclose(res["lp"][:, 0], lp0, rtol=1e-13)
    for s in range(n_steps):
        assert np.allclose(out[:, s + 1], traj[s], rtol=1e-9, atol=1e-12), "diverged at step %d" % s
        if (s + 1) % kw["swp_freq"] != 0:  # (the reference logs ll/lp before the swap of a round's last step)
            assert np.allclose(res["ll"][:, s + 1], lls[s], rtol=1e-9)
            assert np.allclose(res["lp"][:, s + 1], lps[s], rtol=1e-9, atol=1e-9)
    ct, widths = g.counters()
    rc = res["counters"]
    for name in ("step_accept", "step_reject", "swap_accept", "swap_reject", "gauss_accept", "gauss_reject", "de_accept", "de_reject",
                 "fisher_accept", "fisher_reject"):
        assert np.array_equal(ct[name], rc[name]), name
    P = wl.P
    assert np.allclose(widths[:, :P + 1], res["widths"][:, :P + 1], rtol=1e-14) and np.allclose(widths[:, P + 2], res["widths"][:, P + 2], rtol=1e-14)
    assert (ct["step_accept"] + ct["step_reject"] == n_steps).all() and ct["swap_accept"].sum() > 0
    if fisher:
        assert ct["de_accept"].sum() + ct["de_reject"].sum() > 0 and ct["fisher_accept"].sum() + ct["fisher_reject"].sum() > 0
        # the device refreshed exactly when the reference did: once per chain at creation, then at the scheduled steps
        assert (ct["fisher_updates"] + ct["fisher_nan"]).sum() == C + len(refreshes)
    g.close()


def test_log_prior_batch_vs_reference(ctx, oracle):
    """The device priors against the reference's compiled classes (src/standardPriorLibrary.cpp) for the four BASELINE families."""
    from oracle import ref_sampler as rs
    if not hasattr(oracle.lib(), "oracle_ref_log_prior_batch"):
        pytest.skip("oracle/_ref was built without the sampler translation units")
    for cfg in (1, 2, 4, 5):
        wl = workloads.make(cfg, W=512, L=1024 if cfg != 5 else 4096)
        ctx.set_network(wl.detectors, wl.f, wl.psd)
        prior = smp.prior_for(wl)
        rng = np.random.default_rng(100 + cfg)
        params = wl.params.copy()
        k = rng.integers(0, wl.P, 170)
        params[np.arange(170), k] += rng.choice([-1, 1], 170) * 10.0
        got = smp.log_prior_batch(ctx, wl.method, params, prior, wl.mod)
        base = (15 if "Pv2" in wl.method else 11) + (1 if "NRT" in wl.method else 0)
        want = rs.log_prior_batch(wl.method, params, prior, wl.P - base)
        assert np.array_equal(np.isneginf(got), np.isneginf(want))
        fin = ~np.isneginf(want)
        assert 100 < fin.sum() < 512
        assert np.allclose(got[fin], want[fin], rtol=1e-13, atol=1e-13)


def test_lane_split_is_bitwise_invisible(ctx):
    wl = _inject(ctx, workloads.make(2, W=64, L=2048))
    temps = _ladder(4, 8)
    init = _start(wl, 32)
    prior = smp.prior_for(wl)
    out = []
    for lanes in (1, 2):
        s = smp.Sampler(ctx, wl.method, temps, init, prior, wl.gmst, wl.T_segment, wl.mod, seed=5, lanes=lanes, swp_freq=2,
                        history_length=20, fisher_update_number=10, record_cold=1, fisher_deferred=1)
        s.run(37)
        s.run(23)
        out.append((s.state(), s.counters(), s.cold_chains()))
        s.close()
    (a, b) = out
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x, y)
    for k in a[1][0]:
        assert np.array_equal(a[1][0][k], b[1][0][k])
    assert np.array_equal(a[1][1], b[1][1])
    assert a[2].shape == (60, 4, wl.P) and np.array_equal(a[2], b[2])
    # the recorded cold chains end where the state says they are
    assert np.array_equal(a[2][-1], a[0][0][temps == 1.0])


def test_posterior_recovery_at_scale(ctx):
    """64 chains x 3000 steps on a loud two-detector IMRPhenomD injection: the cold chains concentrate on the injection with the
    width the Fisher matrix predicts for the best-measured parameter (chirp mass), proposals of every kind are accepted."""
    wl = _inject(ctx, workloads.make(1, W=64, L=4096))
    temps = _ladder(8, 8, 50.0)
    init = _start(wl, 64, spread=0.05)
    prior = smp.prior_for(wl)
    s = smp.Sampler(ctx, wl.method, temps, init, prior, wl.gmst, wl.T_segment, wl.mod, seed=2026, record_cold=1, history_length=200)
    s.run(3000)
    cold = s.cold_chains(first_step=1500)          # [1500][8][11]
    ct, widths = s.counters()
    pos, ll, lp = s.state()
    assert np.isfinite(ll).all() and np.isfinite(lp).all()
    lnMc = cold[:, :, 7].ravel()
    F, vals, vecs = smp.mcmc_fisher_batch(ctx, wl.method, wl.inj[None, :], wl.gmst, order=4, mod=wl.mod)
    sigma_f = math.sqrt(np.linalg.inv(F[0])[7, 7])
    assert abs(lnMc.mean() - wl.inj[7]) < 5 * sigma_f + 3 * lnMc.std() / math.sqrt(40)
    assert 0.3 * sigma_f < lnMc.std() < 3.0 * sigma_f
    for kind in ("gauss", "de", "fisher", "swap"):
        assert ct[kind + "_accept"].sum() > 0, kind
    assert (ct["step_accept"] + ct["step_reject"] == 3000).all()
    cold_ll = ll[temps == 1.0]
    hot_ll = ll[temps == temps.max()]
    assert cold_ll.mean() > hot_ll.mean()


def test_bad_arguments_fail_loudly(ctx):
    wl = _inject(ctx, workloads.make(1, W=8, L=512))
    prior = smp.prior_for(wl)
    init = _start(wl, 4)
    with pytest.raises(smp.GwatB200Error):
        smp.Sampler(ctx, "IMRPhenomXYZ", np.ones(4), init, prior, wl.gmst, wl.T_segment)
    bad = init.copy()
    bad[2, 8] = 0.3  # eta outside its range: zero prior
    with pytest.raises(smp.GwatB200Error):
        smp.Sampler(ctx, wl.method, np.ones(4), bad, prior, wl.gmst, wl.T_segment)
    with pytest.raises(smp.GwatB200Error):
        smp.Sampler(ctx, wl.method, np.ones(4), init[:, :10], prior, wl.gmst, wl.T_segment)


@pytest.mark.gpu
@pytest.mark.parametrize("n,ens", [(2, 1), (3, 1), (17, 1), (4096, 8), (32768, 8), (36000, 8), (36001, 1), (100003, 7)])
def test_device_swap_sweep_equals_host_sweep(n, ens):
    """The run/pointer-doubling form of chain_swap's sweep (src/mcmc_sampler_internals.cpp:1086-1184) that the sampler runs on the
    device makes the very decisions of the sequential one: random ladders (ensembles of `ens` rungs laid end to end), for the
    shared-memory and the global-scratch sizes, and with the sequential fallback forced."""
    from gw_analysis_tools_b200 import sampler as smp
    from gw_analysis_tools_b200.engine import Context
    rng = np.random.default_rng(n + ens)
    ctx = Context(0)
    temps = np.tile(np.geomspace(1.0, 50.0, ens), n // ens + 1)[:n] if ens > 1 else np.geomspace(1.0, 1e3, n)
    for trial, scale in enumerate([1.0, 30.0, 1e-3]):
        ll = 1e4 + scale * rng.standard_normal(n)
        src_h, acc_h = smp.swap_sweep_host(ll, temps, 1234 + trial, 7 + trial)
        for mode in (0, 1):
            src_d, acc_d = smp.swap_sweep_device(ctx, ll, temps, 1234 + trial, 7 + trial, mode)
            assert np.array_equal(src_d, src_h) and np.array_equal(acc_d, acc_h)
        assert sorted(src_h.tolist()) == list(range(n))  # a permutation


@pytest.mark.gpu
def test_device_swap_sweep_long_runs_fall_back():
    """A state far below every other on a single ladder of increasing temperatures is carried to the top: one run of n - 1 pairs,
    longer than the parallel scan follows -- the CTA falls back to the sequential walk and still agrees."""
    from gw_analysis_tools_b200 import sampler as smp
    from gw_analysis_tools_b200.engine import Context
    ctx = Context(0)
    for n in (300, 5000, 40000):
        temps = np.geomspace(1.0, 1e6, n)
        ll = 1e4 + np.random.default_rng(n).standard_normal(n)
        ll[0] = -1e12
        src_h, acc_h = smp.swap_sweep_host(ll, temps, 99, 3)
        assert acc_h.sum() == n - 1 and src_h[n - 1] == 0
        src_d, acc_d = smp.swap_sweep_device(ctx, ll, temps, 99, 3, 0)
        assert np.array_equal(src_d, src_h) and np.array_equal(acc_d, acc_h)
        # ... and a ladder whose runs stop just short of / just past the cap
        ll2 = ll.copy()
        ll2[0] = 1e4
        ll2[n // 2] = -1e12
        src_h, acc_h = smp.swap_sweep_host(ll2, temps, 99, 3)
        src_d, acc_d = smp.swap_sweep_device(ctx, ll2, temps, 99, 3, 0)
        assert np.array_equal(src_d, src_h) and np.array_equal(acc_d, acc_h)
