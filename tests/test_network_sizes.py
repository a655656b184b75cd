"""Networks of 1, 4 and 5 detectors (the BASELINE configs use 2 and 3): the likelihood and Fisher kernels are instantiated per detector
count, so every count is compared with the compiled reference; larger networks are refused, never truncated."""
import numpy as np
import pytest

from gw_analysis_tools_b200 import abi, engine, workloads

ALL = ["Hanford", "Livingston", "Virgo", "Kagra", "Indigo", "CE", "ET1"]


def _workload(cfg, D, W=24, L=3000):
    wl = workloads.make(cfg, W=W, L=L)
    wl.detectors = ALL[:D]
    # a different noise level per detector, so that a detector mix-up cannot cancel
    wl.psd = workloads.aligo_analytic_psd(wl.f)[None, :] * (1.0 + 0.4 * np.arange(D))[:, None]
    return wl


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [1, 2, 5])
@pytest.mark.parametrize("D", [1, 4, 5])
def test_loglike_vs_oracle_for_every_network_size(ctx, oracle, cfg, D):
    wl = _workload(cfg, D)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    data = 0.9 * np.exp(0.3j) * ctx.coherent_response_batch(wl.method, src)[0]  # a mismatched template
    ref_data = 0.9 * np.exp(0.3j) * oracle.coherent_response(wl.method, src[0], wl.detectors, wl.f)
    assert np.abs(data - ref_data).max() <= 1e-10 * np.abs(ref_data).max()
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    ref = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, data)
    assert np.all(np.isfinite(ref))
    assert (np.abs(got - ref) / np.abs(ref)).max() <= 1e-9, (cfg, D)


@pytest.mark.gpu
@pytest.mark.parametrize("D", [1, 4, 5])
def test_fisher_vs_oracle_for_every_network_size(ctx, oracle, D):
    import fisher_noise
    f = 20.0 + 0.25 * np.arange(2048)
    psd = workloads.aligo_analytic_psd(f)[None, :] * (1.0 + 0.4 * np.arange(D))[:, None]
    dets = ALL[:D]
    ctx.set_network(dets, f, psd)
    srcs = workloads.fisher_sources(4, seed=31)
    got = ctx.fisher_numerical_batch("IMRPhenomD", srcs, 11, order=4, detector_index=-1)  # all detectors summed: one fused launch
    ref = sum(oracle.fisher_numerical_batch("IMRPhenomD", srcs, dets, f, psd, 11, order=4, detector_index=d) for d in range(D))
    floor = sum(fisher_noise.reference_self_difference(oracle, "IMRPhenomD", srcs, dets, f, psd, 11, 4, detector_index=d, runs=2) for d in range(D))
    ok = np.isfinite(ref).all(axis=(1, 2))
    assert ok.sum() >= 3 and np.array_equal(np.isfinite(got).all(axis=(1, 2)), ok)
    err = fisher_noise.normalised_error(got[ok], ref[ok]).reshape(ok.sum(), -1).max(axis=1)
    assert np.all(err <= np.maximum(1e-6, fisher_noise.FACTOR * floor[ok])), (D, err, floor[ok])


@pytest.mark.gpu
def test_networks_beyond_five_detectors_are_refused(ctx):
    f = 20.0 + 0.25 * np.arange(1024)
    for D in (6, 7):
        ctx.set_network(ALL[:D], f, np.tile(workloads.aligo_analytic_psd(f), (D, 1)), np.zeros((D, f.size), dtype=complex))
        wl = workloads.make(1, W=4, L=1024)
        with pytest.raises(engine.GwatB200Error):
            ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
        with pytest.raises(engine.GwatB200Error) as e:
            ctx.fisher_numerical_batch("IMRPhenomD", workloads.fisher_sources(2), 11, order=2, detector_index=-1)
        assert e.value.code == abi.ERR_UNSUPPORTED
        # ... while one detector of the big network at a time is fine
        F = ctx.fisher_numerical_batch("IMRPhenomD", workloads.fisher_sources(2), 11, order=2, detector_index=D - 1)
        assert F.shape == (2, 11, 11)


@pytest.mark.parametrize("name,method", [("D_bbh", "IMRPhenomD"), ("P_full", "IMRPhenomPv2")])
@pytest.mark.parametrize("D", [1, 4, 5])
def test_likelihood_math_for_every_network_size(oracle, name, method, D):
    """CPU tier: the fused per-bin likelihood of gwat_like.h compiled as C++ (tests/host_harness.cpp), instantiated for 1, 4 and 5
    detectors like the kernel, against the compiled reference."""
    import ctypes as C
    import os

    import cases
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hh = C.CDLL(os.path.join(root, "tests", "_build", "libgwat_host_harness.so"))
    gold = np.load(os.path.join(root, "tests", "golden", "waveforms_v1.npz"))
    f = cases.grid([c for c in cases.CASES if c[0] == name][0][3])
    src = cases.source_from_bytes(gold[name + "/src"])
    dets = ALL[:D]
    psd = np.ascontiguousarray(workloads.aligo_analytic_psd(f)[None, :] * (1.0 + 0.4 * np.arange(D))[:, None])
    tmpl = cases.source_from_bytes(gold[name + "/src"])
    tmpl.mass1 *= 1.0004  # the data are another source's responses: a mismatched template
    data = 0.9 * np.exp(0.3j) * oracle.coherent_response(method, tmpl, dets, f)
    ref = oracle.loglike_batch(method, [src], dets, f, psd, data)[0]
    dre, dim = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    names = (C.c_char_p * D)(*[d.encode() for d in dets])
    out = np.zeros(1)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert hh.hh_loglike(method.encode(), 1, C.byref(src), D, names, p(f), f.size, p(psd), p(dre), p(dim), None, 0, 0, 0, p(out)) == 0
    assert abs(out[0] - ref) <= 1e-9 * abs(ref), (D, out[0], ref)
