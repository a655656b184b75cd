"""CPU tier: the sampler mathematics against the reference's OWN compiled code.

oracle/_ref/libgwat_ref.so now also holds src/mcmc_sampler_internals.cpp, src/mcmc_sampler.cpp and
src/standardPriorLibrary.cpp of the reference, compiled unmodified (oracle/Makefile) against stand-in headers for Eigen,
GSL's generators (scripted) and BayesShip, and driven by oracle/sampler_driver.cpp / oracle/ref_sampler.py.  Here:

* the reference's prior classes logPriorStandard_{D,P,D_NRT,P_NRT}[_mod] against the kernels' prior code (host harness build)
  and against the restatement oracle/ptmcmc_ref.py;
* the Eigen stand-in against LAPACK;
* whole runs of the reference's loop PTMCMC_MH_step_incremental -> mcmc_step / gaussian_step / diff_ev_step / fisher_step /
  update_fisher / update_history / update_step_widths / chain_swap, fed with the counter-based draws of the CUDA sampler,
  against the restatement driven by the same draws and the same compiled likelihood: positions, likelihoods, priors, widths and
  counters agree step for step.  This pins the restatement (which the GPU tier also uses) to the reference itself; the GPU tier
  (tests/test_sampler_gpu.py::test_trajectories_match_reference_steps) compares the device with the compiled reference directly.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from gw_analysis_tools_b200 import abi, workloads
from oracle import ptmcmc_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


from test_sampler_host import PRIOR, _positions, c_prior  # noqa: E402


@pytest.fixture(scope="module")
def rs(oracle):
    from oracle import ref_sampler
    if not hasattr(oracle.lib(), "oracle_sampler_create"):
        pytest.skip("oracle/_ref was built without the sampler translation units")
    return ref_sampler


@pytest.fixture(scope="module")
def hh():
    path = os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so")
    lib = C.CDLL(path)
    lib.hh_log_prior.restype = C.c_double
    return lib


@pytest.mark.parametrize("pv2,nrt,nmod,tidal_love", [(0, 0, 0, 1), (1, 0, 0, 1), (0, 0, 2, 1), (1, 0, 1, 1), (0, 1, 0, 1), (0, 1, 1, 1), (1, 1, 0, 1),
                                                     (0, 1, 0, 0), (1, 1, 0, 0), (1, 1, 2, 0)])
def test_reference_priors(rs, hh, pv2, nrt, nmod, tidal_love):
    """The compiled reference classes == the kernels' prior code == the restatement, on 4000 scattered points per family."""
    rng = np.random.default_rng(10 * pv2 + nrt + 100 * nmod + 7)
    pd = dict(PRIOR, tidal_love=tidal_love)
    cp = c_prior(pd)
    pos = _positions(rng, 4000, pv2, nrt * (1 if tidal_love else 2), nmod)
    method = ("IMRPhenomPv2" if pv2 else "IMRPhenomD") + ("_NRT" if nrt else "")
    want = rs.log_prior_batch(method, pos, cp, nmod)
    got = np.array([hh.hh_log_prior(C.byref(cp), pv2, nrt, pos.shape[1], _p(row)) for row in pos])
    rst = np.array([ref.standard_log_prior(list(row), pd, bool(pv2), bool(nrt)) for row in pos])
    assert np.array_equal(np.isneginf(want), np.isneginf(got)) and np.array_equal(np.isneginf(want), np.isneginf(rst))
    fin = ~np.isneginf(want)
    if pv2 and nrt and tidal_love:
        assert fin.sum() == 0  # the reference's binary-Love test reads pos[11] = cos(tilt_1) here (standardPriorLibrary.cpp:499): nothing passes
    else:
        assert 20 < fin.sum() < 3900
        assert np.abs(got[fin] - want[fin]).max() <= 1e-13 * np.maximum(1, np.abs(want[fin])).max()
        assert np.abs(rst[fin] - want[fin]).max() <= 1e-13 * np.maximum(1, np.abs(want[fin])).max()


def test_eigen_standin_against_lapack(rs):
    rng = np.random.default_rng(3)
    for n in (2, 7, 11, 15):
        A = rng.standard_normal((n, n))
        A = A @ A.T + np.diag(rng.uniform(0, 1e3, n))
        vals, vecs = rs.eigen_standin(A)
        w = np.linalg.eigvalsh(A)
        assert np.allclose(vals, w, rtol=1e-12, atol=1e-12 * np.abs(w).max())
        assert np.allclose(vecs @ vecs.T, np.eye(n), atol=1e-12)
        assert np.allclose(vecs @ A @ vecs.T, np.diag(vals), atol=1e-9 * np.abs(w).max())
        assert (vecs[np.arange(n), np.abs(vecs).argmax(axis=1)] > 0).all()  # sign convention of the stand-in


def _workload(oracle, cfg, L):
    wl = workloads.make(cfg, W=64, L=L)
    _, src = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.inj[None, :], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, None,
                                       return_sources=True)
    wl.data = oracle.coherent_response(wl.method, src[0], wl.detectors, wl.f)
    return wl


def _prior_for(wl):
    ns = "NRT" in wl.method
    d = dict(PRIOR, mass1_prior=[.5, 3.] if ns else [1., 100.], mass2_prior=[.5, 3.] if ns else [1., 100.],
             spin1_prior=[-.05, .05] if ns else [-.95, .95], spin2_prior=[-.05, .05] if ns else [-.95, .95], a1_prior=[0, .95], a2_prior=[0, .95],
             DL_prior=[1., 10000.], T_merger=float(wl.inj[5]), mod_priors=[[0., 50.]] * abi.MAX_MOD)
    return d, c_prior(d)


@pytest.mark.parametrize("cfg,fisher", [(1, False), (1, True), (2, True), (4, True)])
def test_reference_loop_matches_restatement(oracle, rs, cfg, fisher):
    """The reference's compiled sampler loop and the restatement, same draws, same compiled likelihood: identical runs."""
    wl = _workload(oracle, cfg, 512)
    Cn = 8
    temps = np.tile(np.geomspace(1.0, 20.0, 4), 2)
    init = wl.inj[None, :] + 0.2 * (wl.params[:Cn] - wl.inj[None, :])
    pd, cp = _prior_for(wl)
    pv2, nrt = "Pv2" in wl.method, "NRT" in wl.method
    kw = dict(swp_freq=3, history_length=12, history_update=2, fisher_update_number=4, check_stepsize_freq=5)
    seed = 31 + cfg
    n_rounds = 14
    n_steps = n_rounds * kw["swp_freq"]
    P = wl.P

    def ll(p):
        return oracle.loglike_mcmc_batch(wl.method, wl.mod, np.atleast_2d(p), wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data)

    rng = np.random.default_rng(cfg)
    systems = {}

    def fish(c, p):
        # any eigen-system will do for this comparison, as long as both sides jump along the same one: a random rotation with a
        # spread of eigenvalues on both sides of the |lambda| < 10 clamp of fisher_step
        q, _ = np.linalg.qr(rng.standard_normal((P, P)))
        vals = np.sort(10 ** rng.uniform(-1, 6, P))
        systems[(r_box[0].step if r_box else -1, c)] = (vals, q)
        return vals, q
    r_box = []
    r = ref.Sampler(ll, lambda p: ref.standard_log_prior(list(p), pd, pv2, nrt), temps, init, seed, fisher=fish if fisher else None, **kw)
    r_box.append(r)
    traj, lls, lps = [], [], []
    for _ in range(n_steps):
        r.run(1)
        traj.append(np.array(r.pos))
        lls.append(np.array(r.ll))
        lps.append(np.array(r.lp))
    R = rs.RefSampler(wl, temps, init, cp, seed, n_rounds, ref.uniform2, ref.normal_from, fisher_exist=fisher,
                      initial_fisher=(lambda c: systems[(-1, c)]) if fisher else None, **kw)
    kinds, refreshes = R.script(lambda s, c: systems[(s, c)])
    assert sorted(refreshes) == sorted(k for k in systems if k[0] >= 0)
    R.run()
    res = R.results()
    R.close()
    d = res["diag"]
    assert d["rng_underflow"] == 0 and d["uniforms_left"] == 0 and d["normals_left"] == 0 and d["fisher_script_underflow"] == 0, d
    assert (res["chain_pos"] == n_steps).all()
    out = res["output"]
    assert np.array_equal(out[:, 0], init)
    for s in range(n_steps):
        assert np.allclose(out[:, s + 1], traj[s], rtol=1e-13, atol=1e-15), "positions part at step %d" % s
        boundary = (s + 1) % kw["swp_freq"] == 0
        if not boundary:  # (the reference logs ll/lp before the swap of a round's last step)
            assert np.allclose(res["ll"][:, s + 1], lls[s], rtol=1e-13)
            assert np.allclose(res["lp"][:, s + 1], lps[s], rtol=1e-13, atol=1e-13)
    ct = res["counters"]
    assert np.array_equal(ct["step_accept"], [c["step"][0] for c in r.ct]) and np.array_equal(ct["step_reject"], [c["step"][1] for c in r.ct])
    assert np.array_equal(ct["swap_accept"], [c["swap"][0] for c in r.ct]) and np.array_equal(ct["swap_reject"], [c["swap"][1] for c in r.ct])
    for name in ("gauss", "de", "fisher"):
        assert np.array_equal(ct[name + "_accept"], [c[name][0] for c in r.ct]), name
        assert np.array_equal(ct[name + "_reject"], [c[name][1] for c in r.ct]), name
    assert np.allclose(res["widths"][:, :P], np.array(r.widths)[:, :P], rtol=1e-14)
    assert np.allclose(res["widths"][:, P], np.array(r.widths)[:, P], rtol=1e-14) and np.allclose(res["widths"][:, P + 2], np.array(r.widths)[:, P + 2], rtol=1e-14)
    assert ct["swap_accept"].sum() > 0 and ct["step_accept"].sum() > 0
    if fisher:
        assert ct["fisher_accept"].sum() + ct["fisher_reject"].sum() > 0 and ct["de_accept"].sum() + ct["de_reject"].sum() > 0
        assert len(refreshes) >= Cn // 2
        # the eigen-systems the reference ended up with are the scripted ones (update_fisher's storage layout, :683-692)
        last = {c: systems[(-1, c)] for c in range(Cn)}
        for (s, c) in sorted(refreshes):
            last[c] = systems[(s, c)]
        for c, (vals, vecs) in last.items():
            assert np.allclose(res["fvals"][c], vals, rtol=1e-14) and np.allclose(res["fvecs"][c], vecs, rtol=1e-14, atol=1e-300)
