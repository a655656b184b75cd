"""GPU parity: the CUDA path through the C ABI against the oracle (the reference's own code) on identical inputs.

Tolerances are BASELINE.json's: waveform <= 1e-10 relative to max|h|, log-likelihood <= 1e-9 relative.
"""
import numpy as np
import pytest

from gw_analysis_tools_b200 import abi, workloads

pytestmark = pytest.mark.gpu

WF_TOL = 1e-10
LL_TOL = 1e-9


def _sources_from_oracle(oracle, wl, n):
    _, srcs = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params[:n], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd,
                                        None, return_sources=True)
    return srcs


def _inject(oracle, wl):
    _, src = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.inj[None, :], wl.gmst, wl.T_segment, wl.detectors, wl.f,
                                       wl.psd, None, return_sources=True)
    wl.data = oracle.coherent_response(wl.method, src[0], wl.detectors, wl.f)
    return wl


def _relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("masses", [(36.0, 29.0), (10.0, 8.0)])
def test_phenomd_waveform_and_response(ctx, oracle, masses):
    wl = workloads.make(1, W=8, L=4096, masses=masses)
    srcs = _sources_from_oracle(oracle, wl, 8)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    hp, hc = ctx.fourier_waveform_batch(wl.method, srcs)
    resp = ctx.coherent_response_batch(wl.method, srcs)
    for w in range(8):
        rp, rc = oracle.fourier_waveform(wl.method, srcs[w], wl.f)
        assert _relerr(hp[w], rp) <= WF_TOL
        assert _relerr(hc[w], rc) <= WF_TOL
        rr = oracle.coherent_response(wl.method, srcs[w], wl.detectors, wl.f)
        for d in range(wl.D):
            assert _relerr(resp[w, d], rr[d]) <= WF_TOL


@pytest.mark.parametrize("masses", [(36.0, 29.0), (10.0, 8.0)])
def test_phenomd_loglike_mcmc(ctx, oracle, masses):
    wl = _inject(oracle, workloads.make(1, W=64, L=8192, masses=masses))
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment)
    ref = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data)
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= LL_TOL, rel.max()
    assert ctx.last_active_bins > 0


def test_phenompv2_waveform_and_response(ctx, oracle):
    wl = workloads.make(2, W=8, L=4096)
    srcs = _sources_from_oracle(oracle, wl, 8)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    hp, hc = ctx.fourier_waveform_batch(wl.method, srcs)
    resp = ctx.coherent_response_batch(wl.method, srcs)
    for w in range(8):
        rp, rc = oracle.fourier_waveform(wl.method, srcs[w], wl.f)
        assert _relerr(hp[w], rp) <= WF_TOL
        assert _relerr(hc[w], rc) <= WF_TOL
        rr = oracle.coherent_response(wl.method, srcs[w], wl.detectors, wl.f)
        for d in range(wl.D):
            assert _relerr(resp[w, d], rr[d]) <= WF_TOL


def test_phenompv2_loglike_mcmc(ctx, oracle):
    wl = _inject(oracle, workloads.make(2, W=64, L=8192))
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment)
    ref = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data)
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= LL_TOL, rel.max()


def test_repack_matches_reference(ctx, oracle):
    for cfg in (1, 2):
        wl = workloads.make(cfg, W=16, L=1024)
        srcs = _sources_from_oracle(oracle, wl, 16)
        ctx.set_network(wl.detectors, wl.f, wl.psd)
        mine = ctx.repack_mcmc_batch(wl.method, wl.params[:16], wl.gmst, wl.mod)
        for a, b in zip(mine, srcs):
            for name in ("mass1", "mass2", "Luminosity_Distance", "RA", "DEC", "psi", "incl_angle", "phiRef"):
                assert abs(getattr(a, name) - getattr(b, name)) <= 4e-16 * max(1.0, abs(getattr(b, name))), name
            assert abs((wl.T_segment - a.tc) - b.tc) <= 1e-15 * wl.T_segment
            # the aligned-spin repack of the reference leaves the in-plane components uninitialised (src/fisher.cpp:2276-2277)
            for i in (range(3) if cfg == 2 else [2]):
                assert abs(a.spin1[i] - b.spin1[i]) <= 1e-15 and abs(a.spin2[i] - b.spin2[i]) <= 1e-15


def test_antenna_and_dtoa(ctx, oracle):
    rng = np.random.default_rng(5)
    W = 257
    RA, DEC, psi = rng.uniform(0, 2 * np.pi, W), np.arcsin(rng.uniform(-1, 1, W)), rng.uniform(0, np.pi, W)
    dets = ["Hanford", "Livingston", "Virgo", "Kagra", "Indigo", "CE", "ET1", "ET2"]
    f = 20 + np.arange(64.0)
    ctx.set_network(dets, f, np.ones((len(dets), 64)))
    fp, fc, dt = ctx.antenna_batch(RA, DEC, psi, 2.1)
    rfp, rfc, rdt = oracle.antenna_batch(RA, DEC, psi, 2.1, dets)
    assert np.abs(fp - rfp).max() <= 1e-14 and np.abs(fc - rfc).max() <= 1e-14
    assert np.abs(dt - rdt).max() <= 1e-16


def test_error_paths(ctx):
    from gw_analysis_tools_b200 import engine
    f = 20 + np.arange(64.0)
    with pytest.raises(engine.GwatB200Error):
        ctx.set_network(["Atlantis"], f, np.ones((1, 64)))
    ctx.set_network(["Hanford"], f, np.ones((1, 64)))
    with pytest.raises(engine.GwatB200Error):  # no data uploaded
        ctx.loglike_mcmc_batch("IMRPhenomD", np.zeros((1, 11)), 0.0, 1.0)
    ctx.set_network(["Hanford"], f, np.ones((1, 64)), np.zeros((1, 64), complex))
    with pytest.raises(engine.GwatB200Error):  # unknown method
        ctx.loglike_mcmc_batch("IMRPhenomXYZ", np.zeros((1, 11)), 0.0, 1.0)
    with pytest.raises(engine.GwatB200Error):  # wrong dimension
        ctx.loglike_mcmc_batch("IMRPhenomD", np.zeros((1, 12)), 0.0, 1.0)
    # an unphysical point (eta > 1/4) gives NaN, like the reference, and does not poison its neighbours
    wl = workloads.make(1, W=4, L=64)
    bad = wl.params.copy()
    bad[1, 8] = 0.3
    ctx.set_network(wl.detectors, wl.f, wl.psd, np.zeros((2, 64), complex))
    out = ctx.loglike_mcmc_batch(wl.method, bad, wl.gmst, wl.T_segment)
    assert np.isnan(out[1]) and np.all(np.isfinite(out[[0, 2, 3]]))
    assert ctx.loglike_mcmc_batch(wl.method, bad[:0], wl.gmst, wl.T_segment).size == 0
