"""The reference's INTRINSIC sampling mode (mcmc_intrinsic: PTMCMC_method_specific_prep, src/mcmc_gw.cpp:1880-1985): the chain samples
ln Mc, eta and the spins only (4 parameters for the IMRPhenomD family, + tidal for NRT; 8 for IMRPhenomPv2; then the modifications),
the likelihood is maximised over tc and phic (:2611-2722) and the Fisher matrix is the sky-averaged one of the "MCMC_" set
(src/fisher.cpp:183-338, 2000-2013, 2308-2376) with the intrinsic branch of MCMC_fisher_transformations (src/mcmc_gw.cpp:2163-2179).

CPU tier: the GWAT_HD repack and Fisher mathematics compiled as C++ against the reference build.  GPU tier: the C ABI against it."""
import ctypes as C
import os

import numpy as np
import pytest

import fisher_noise
import test_fisher_sky as tfs
from gw_analysis_tools_b200 import abi, engine, sampler, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)
GMST = 2.1


def repack_cases():
    out = []
    out.append(("IMRPhenomD", abi.mod_defaults(), [np.log(20.), 0.24, 0.3, -0.2]))
    m = abi.mod_defaults(ppE_Nmod=1, bppe=[-1.0])
    out.append(("ppE_IMRPhenomD_Inspiral", m, [np.log(15.), 0.2, -0.5, 0.1, 0.03]))
    out.append(("dCS_IMRPhenomD", m, [np.log(15.), 0.2, -0.5, 0.1, 12.0]))  # sqrt(alpha) in km -> alpha^2 in s^4 (src/mcmc_gw.cpp:2560-2565)
    m = abi.mod_defaults(gIMR_Nmod_phi=1, gIMR_phii=[4], gIMR_Nmod_beta=1, gIMR_betai=[2])
    out.append(("gIMRPhenomD", m, [np.log(15.), 0.2, -0.5, 0.1, 0.05, 0.02]))
    out.append(("IMRPhenomD_NRT", abi.mod_defaults(tidal_love=1), [np.log(1.2), 0.249, 0.01, -0.02, np.log(400.)]))
    out.append(("IMRPhenomD_NRT", abi.mod_defaults(tidal_love=0), [np.log(1.2), 0.249, 0.01, -0.02, np.log(400.), np.log(300.)]))
    out.append(("IMRPhenomPv2", abi.mod_defaults(), [np.log(25.), 0.22, 0.6, 0.4, 0.3, -0.5, 1.0, 4.0]))
    out.append(("IMRPhenomPv2", abi.mod_defaults(), [np.log(25.), 0.22, 0.6, 0.4, 1.0000001, -1.2, 1.0, 4.0]))  # cosines outside [-1, 1] are clamped
    return [(a, b, np.array(c)) for a, b, c in out]


# members the reference's repack leaves uninitialised for a set (src/fisher.cpp:2308-2376): not compared
UNSET = {"theta", "phi", "theta_l", "phi_l"}


def _same_record(got, ref, aligned):
    for name, _ in abi.Source._fields_:
        if name in UNSET or name.startswith("reserved"):
            continue
        a, b = getattr(got, name), getattr(ref, name)
        if hasattr(a, "__len__"):
            idx = [2] if (aligned and name in ("spin1", "spin2")) else range(len(a))  # aligned sets never write the in-plane components
            for i in idx:
                assert abs(a[i] - b[i]) <= 1e-15 * max(1.0, abs(b[i])), (name, i, a[i], b[i])
        else:
            assert a == b or abs(a - b) <= 4e-16 * max(1.0, abs(b)), (name, a, b)


@pytest.mark.parametrize("k", range(len(repack_cases())))
def test_intrinsic_repack_math_vs_reference(oracle, k):
    hh = C.CDLL(os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so"))
    method, mod, par = repack_cases()[k]
    ref = oracle.repack_mcmc_intrinsic(method, mod, par[None, :], GMST)[0]
    got = abi.Source()
    assert hh.hh_repack_mcmc_intrinsic(method.encode(), C.byref(mod), par.size, par.ctypes.data_as(_dp), C.c_double(GMST), C.byref(got)) == 0
    assert got.sky_average == 1 and ref.sky_average == 1
    _same_record(got, ref, aligned="Pv2" not in method)


def fisher_cases():
    """(method, dimension, sources): sky-averaged records of the intrinsic sets, plain and modified families."""
    f, psd, srcs = tfs.setup_case()
    out = [("IMRPhenomD", 4, srcs[:4])]
    for method, dim, ss in tfs.mod_cases()[2]:
        out.append((method, dim - 3, ss))
    return f, psd, out


@pytest.mark.parametrize("order", [2, 4])
def test_intrinsic_fisher_math_vs_reference(oracle, order):
    hh = C.CDLL(os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so"))
    f, psd, cs = fisher_cases()
    for method, dim, srcs in cs:
        m = "MCMC_" + method
        ref = oracle.fisher_numerical_batch(m, srcs, ["Hanford"], f, psd[None, :], dim, order=order, detector_index=0)
        got = np.zeros_like(ref)
        for i, s in enumerate(srcs):
            assert hh.hh_fisher_numerical(m.encode(), b"Hanford", b"Hanford", dim, order, C.byref(s), f.ctypes.data_as(_dp), f.size,
                                          psd.ctypes.data_as(_dp), got[i].ctypes.data_as(_dp)) == 0, method
        assert np.all(np.isfinite(ref)) and np.all(np.isfinite(got)), method
        floor = fisher_noise.reference_self_difference(oracle, m, srcs, ["Hanford"], f, psd[None, :], dim, order, detector_index=0, runs=2)
        err = tfs.normalised(got, ref).reshape(len(srcs), -1).max(axis=1)
        assert np.all(err <= np.maximum(tfs.NORM_TOL, fisher_noise.FACTOR * floor)), (method, err, floor)


def pv2_fisher_case(oracle):
    """Four records of the 8-parameter IMRPhenomPv2 set, made by the reference's own repack; two detectors with one PSD."""
    f, psd, _ = tfs.setup_case()
    rng = np.random.default_rng(5)
    pars = np.array([[np.log(rng.uniform(8, 30)), rng.uniform(0.15, 0.249), rng.uniform(.1, .8), rng.uniform(.1, .8), rng.uniform(-.9, .9),
                      rng.uniform(-.9, .9), rng.uniform(0, 6), rng.uniform(0, 6)] for _ in range(4)])
    return f, psd, pars, list(oracle.repack_mcmc_intrinsic("IMRPhenomPv2", abi.mod_defaults(), pars, GMST))


@pytest.mark.parametrize("order,di", [(2, 0), (4, 1)])
def test_intrinsic_pv2_fisher_math_vs_reference(oracle, order, di):
    """A sky-averaged IMRPhenomPv2 record takes the RESPONSE branch of calculate_derivatives (src/fisher.cpp:340-557) with the 8
    intrinsic parameters (unpack :1968-1990, repack :2308-2352: extrinsic members at constants)."""
    hh = C.CDLL(os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so"))
    f, psd, pars, srcs = pv2_fisher_case(oracle)
    dets = ["Hanford", "Livingston"]
    ref = oracle.fisher_numerical_batch("MCMC_IMRPhenomPv2", srcs, dets, f, np.stack([psd, psd]), 8, order=order, detector_index=di)
    got = np.zeros_like(ref)
    for i, s in enumerate(srcs):
        assert hh.hh_fisher_numerical(b"MCMC_IMRPhenomPv2", dets[di].encode(), b"Hanford", 8, order, C.byref(s), f.ctypes.data_as(_dp), f.size,
                                      psd.ctypes.data_as(_dp), got[i].ctypes.data_as(_dp)) == 0
    assert np.all(np.isfinite(ref)) and np.all(np.isfinite(got))
    floor = fisher_noise.reference_self_difference(oracle, "MCMC_IMRPhenomPv2", srcs, dets, f, np.stack([psd, psd]), 8, order, detector_index=di, runs=2)
    err = tfs.normalised(got, ref).reshape(len(srcs), -1).max(axis=1)
    assert np.all(err <= np.maximum(tfs.NORM_TOL, fisher_noise.FACTOR * floor)), (err, floor)


# ---------------------------------------------------------------------------------------------------------------------------------
def _network(ctx, oracle, L=2048):
    """Three detectors, a uniform grid and an injection as data (the maximised likelihoods need all three)."""
    f = 20. + 0.25 * np.arange(L)
    dets = ["Hanford", "Livingston", "Virgo"]
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1)) * np.array([1.0, 1.3, 2.0])[:, None]
    inj = abi.source_defaults(mass1=31., mass2=24., Luminosity_Distance=400., spin1=[0, 0, .2], spin2=[0, 0, -.1], RA=1., DEC=.3, psi=.4,
                              incl_angle=.6, gmst=GMST, f_ref=20., phiRef=1.3, tc=2.5)
    data = oracle.coherent_response("IMRPhenomD", inj, dets, f)
    ctx.set_network(dets, f, psd, data)
    return dets, f, psd, data


@pytest.mark.gpu
def test_intrinsic_repack_vs_reference(ctx, oracle):
    for method, mod, par in repack_cases():
        pars = np.tile(par, (5, 1))
        pars[1:, 2:4] += np.random.default_rng(1).uniform(-.05, .05, (4, 2))
        ref = oracle.repack_mcmc_intrinsic(method, mod, pars, GMST)
        got = ctx.repack_mcmc_intrinsic_batch(method, pars, GMST, mod)
        for a, b in zip(got, ref):
            _same_record(a, b, aligned="Pv2" not in method)
    with pytest.raises(engine.GwatB200Error):  # the 11-dimensional vector is not an intrinsic set
        ctx.repack_mcmc_intrinsic_batch("IMRPhenomD", np.zeros((1, 11)), GMST)


@pytest.mark.gpu
@pytest.mark.parametrize("method,k", [("IMRPhenomD", 0), ("ppE_IMRPhenomD_Inspiral", 1), ("IMRPhenomD_NRT", 4), ("IMRPhenomPv2", 6)])
def test_intrinsic_maximized_likelihood_vs_reference(ctx, oracle, method, k):
    dets, f, psd, data = _network(ctx, oracle)
    m, mod, par = repack_cases()[k]
    assert m == method
    rng = np.random.default_rng(4)
    pars = np.tile(par, (6, 1))
    if "NRT" not in method:
        pars[:, 0] = np.log(np.exp(pars[:, 0]) * rng.uniform(0.9, 1.4, 6))
    pars[:, 2:4] += rng.uniform(-.1, .1, (6, 2))
    got = ctx.loglike_maximized_mcmc_batch(method, pars, GMST, mod)
    # the reference's chain: its repack of the same vectors, then its maximised likelihood (src/mcmc_gw.cpp:2611-2722)
    ref = oracle.loglike_maximized_batch(method, list(oracle.repack_mcmc_intrinsic(method, mod, pars, GMST)), dets, f, psd, data)
    assert np.all(np.isfinite(ref)) and len(set(np.round(ref, 6))) == 6
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max(), (got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 4])
def test_intrinsic_fisher_vs_reference(ctx, oracle, order):
    f, psd, cs = fisher_cases()
    ctx.set_network(["Hanford", "Livingston"], f, np.stack([psd, 2.0 * psd]))
    for method, dim, srcs in cs:
        m = "MCMC_" + method
        ref = oracle.fisher_numerical_batch(m, srcs, ["Hanford"], f, psd[None, :], dim, order=order, detector_index=0)
        got = ctx.fisher_numerical_batch(m, srcs, dim, order=order, detector_index=0)
        assert np.all(np.isfinite(got)), method
        floor = fisher_noise.reference_self_difference(oracle, m, srcs, ["Hanford"], f, psd[None, :], dim, order, detector_index=0, runs=2)
        err = tfs.normalised(got, ref).reshape(len(srcs), -1).max(axis=1)
        assert np.all(err <= np.maximum(tfs.NORM_TOL_GPU, fisher_noise.FACTOR * floor)), (method, err, floor)
        assert np.median(tfs.normalised(got, ref)) <= 1e-6, method
        assert np.array_equal(got, np.swapaxes(got, 1, 2))
    # refused with the reason, never computed as something else
    for method, dim in (("IMRPhenomPv2", 13), ("MCMC_IMRPhenomD_NRT", 5)):
        with pytest.raises(engine.GwatB200Error) as e:
            ctx.fisher_numerical_batch(method, cs[0][2], dim, order=order, detector_index=0)
        assert e.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(engine.GwatB200Error) as e:
        ctx.fisher_numerical_batch("MCMC_IMRPhenomD", cs[0][2], 7, order=order, detector_index=0)
    assert e.value.code == abi.ERR_ARG


@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 4])
def test_intrinsic_pv2_fisher_vs_reference(ctx, oracle, order):
    f, psd, pars, srcs = pv2_fisher_case(oracle)
    dets = ["Hanford", "Livingston"]
    ctx.set_network(dets, f, np.stack([psd, psd]))
    for di in (0, 1, -1):
        got = ctx.fisher_numerical_batch("MCMC_IMRPhenomPv2", srcs, 8, order=order, detector_index=di)
        if di >= 0:
            ref = oracle.fisher_numerical_batch("MCMC_IMRPhenomPv2", srcs, dets, f, np.stack([psd, psd]), 8, order=order, detector_index=di)
            floor = fisher_noise.reference_self_difference(oracle, "MCMC_IMRPhenomPv2", srcs, dets, f, np.stack([psd, psd]), 8, order, detector_index=di, runs=2)
        else:
            ref = sum(oracle.fisher_numerical_batch("MCMC_IMRPhenomPv2", srcs, dets, f, np.stack([psd, psd]), 8, order=order, detector_index=d) for d in (0, 1))
        assert np.all(np.isfinite(got))
        err = tfs.normalised(got, ref).reshape(len(srcs), -1).max(axis=1)
        assert np.all(err <= np.maximum(tfs.NORM_TOL_GPU, fisher_noise.FACTOR * floor)), (di, err, floor)
        assert np.median(tfs.normalised(got, ref)) <= 1e-6
    # the wrapper of an intrinsic run: summed over the detectors, prior terms in place of the diagonal of eta, the spins and their angles
    F = sampler.mcmc_fisher_intrinsic_batch(ctx, "IMRPhenomPv2", pars, GMST, order=order)
    d = np.diagonal(F, axis1=1, axis2=2)
    assert np.array_equal(d[:, 1:], np.tile([4.0, .25, .25, .25, .25, 1 / (4 * np.pi ** 2), 1 / (4 * np.pi ** 2)], (4, 1)))
    dg = np.sqrt(np.abs(np.diagonal(ref, axis1=1, axis2=2)))
    keep = ~np.eye(8, dtype=bool)
    keep[0, 0] = True  # everything but the replaced diagonal entries
    assert (np.abs(F - ref) / (dg[:, :, None] * dg[:, None, :]))[:, keep].max() <= 3 * float(np.max(floor)) + tfs.NORM_TOL_GPU


@pytest.mark.gpu
def test_intrinsic_mcmc_fisher_wrapper_vs_reference(ctx, oracle):
    """MCMC_fisher_wrapper with mcmc_intrinsic: the reference's pieces (repack, per-detector fisher_numerical, summed) and the intrinsic
    transformations restated from src/mcmc_gw.cpp:2163-2193 -- entries (1,1), (2,2), (3,3) REPLACED by 4, 1/4, 1/4; dCS unit factor."""
    dets, f, psd, data = _network(ctx, oracle, L=1024)
    for method, mod, par in (repack_cases()[0], repack_cases()[2]):
        dim = par.size
        pars = np.tile(par, (3, 1))
        pars[:, 2:4] += np.random.default_rng(8).uniform(-.1, .1, (3, 2))
        got = sampler.mcmc_fisher_intrinsic_batch(ctx, method, pars, GMST, order=4, mod=mod)
        srcs = list(oracle.repack_mcmc_intrinsic(method, mod, pars, GMST))
        ref = sum(oracle.fisher_numerical_batch("MCMC_" + method, srcs, dets, f, psd, dim, order=4, detector_index=d) for d in range(3))
        dg = np.sqrt(np.abs(np.diagonal(ref, axis1=1, axis2=2))).copy()  # the yardstick: the diagonal BEFORE the prior terms replace it
        for w in range(3):
            ref[w, 1, 1], ref[w, 2, 2], ref[w, 3, 3] = 4.0, 0.25, 0.25
            if method.startswith("dCS"):
                factor = 4 * srcs[w].betappe[0] ** 0.75 * 1000 / 299792458.
                ref[w, dim - 1, :] *= factor
                ref[w, :, dim - 1] *= factor
                dg[w, dim - 1] *= factor
        assert np.all(np.isfinite(got))
        assert (np.abs(got - ref) / (dg[:, :, None] * dg[:, None, :])).max() <= tfs.NORM_TOL_GPU, method
        assert got[0, 1, 1] == 4.0 and got[0, 2, 2] == 0.25 and got[0, 3, 3] == 0.25


# ---- sky_average in the waveform path: the prefactor of A0 and nothing else (populate_source_parameters, src/util.cpp:1024-1025) ----------
SKY_WF = (("D_bbh", "IMRPhenomD"), ("P_full", "IMRPhenomPv2"), ("NRT_love", "IMRPhenomD_NRT"))


def _sky_wf_case(name):
    import cases
    gold = np.load(os.path.join(ROOT, "tests", "golden", "waveforms_v1.npz"))
    f = cases.grid([c for c in cases.CASES if c[0] == name][0][3])
    s = cases.source_from_bytes(gold[name + "/src"])
    s.sky_average = 1
    return f, s, gold[name + "/hp"]


@pytest.mark.parametrize("name,method", SKY_WF)
def test_sky_averaged_waveform_math_vs_reference(oracle, name, method):
    hh = C.CDLL(os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so"))
    f, s, hp_pointed = _sky_wf_case(name)
    hp, hc = oracle.fourier_waveform(method, s, f)
    o = [np.zeros(f.size) for _ in range(4)]
    assert hh.hh_fourier_waveform(method.encode(), C.byref(s), f.ctypes.data_as(_dp), f.size, *[x.ctypes.data_as(_dp) for x in o]) == 0
    assert np.abs(o[0] + 1j * o[1] - hp).max() <= 1e-10 * np.abs(hp).max() and np.abs(o[2] + 1j * o[3] - hc).max() <= 1e-10 * np.abs(hc).max()
    # sqrt(pi/30) / sqrt(pi 40/192) = 2/5 of the pointed amplitude
    assert abs(np.abs(hp).max() / np.abs(hp_pointed).max() - 0.4) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name,method", SKY_WF)
def test_sky_averaged_waveform_vs_reference(ctx, oracle, name, method):
    f, s, _ = _sky_wf_case(name)
    ctx.set_network(["Hanford"], f, np.ones((1, f.size)))
    hp, hc = oracle.fourier_waveform(method, s, f)
    g_hp, g_hc = ctx.fourier_waveform_batch(method, [s])
    assert np.abs(g_hp[0] - hp).max() <= 1e-10 * np.abs(hp).max() and np.abs(g_hc[0] - hc).max() <= 1e-10 * np.abs(hc).max()
