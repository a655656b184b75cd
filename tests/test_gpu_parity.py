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
