"""Sky-averaged Fisher matrices (calculate_derivatives, sky-averaged branch, src/fisher.cpp:183-338): IMRPhenomD in the
7-parameter set ln A0, phic, tc, ln Mc, ln eta, chi_s, chi_a, from derivatives of amplitude and phase.

CPU tier: the GWAT_HD mathematics (unpack/repack, amplitude/phase per bin, stencil) compiled as C++ against the reference
build.  GPU tier: gwat_b200_fisher_numerical_batch against the reference build.  Tolerance: the Fisher noise floor of the
reference itself (eps = 1e-8 stencil; its FMA and non-FMA builds differ by 2e-6..5e-6 in |dF_ij|/sqrt(F_ii F_jj)).
"""
import ctypes as C
import os

import numpy as np
import pytest

import fisher_noise
from gw_analysis_tools_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NORM_TOL = 2e-5  # host mathematics (same compiler as the reference): 4 x the upper end of the reference's own noise floor
NORM_TOL_GPU = 1e-4  # device arithmetic contracts FMAs differently: largest element; the median must stay below 1e-6
_dp = C.POINTER(C.c_double)


def setup_case():
    # the configuration of the reference's testing/fisher_comparison.cpp: 3000 bins 15-1000 Hz, Hanford_O1_fitted
    L = 3000
    f = 15 + np.arange(L) * ((1000 - 15.) / (L - 1))
    psd = engine.populate_noise(f, "Hanford_O1_fitted") ** 2
    rng = np.random.default_rng(2)
    srcs = []
    for _ in range(6):
        m = np.sort(rng.uniform(5, 60, 2))[::-1]
        srcs.append(abi.source_defaults(mass1=m[0], mass2=m[1], Luminosity_Distance=rng.uniform(100, 900),
                                        spin1=[0, 0, rng.uniform(-.8, .8)], spin2=[0, 0, rng.uniform(-.8, .8)], tc=rng.uniform(0, 5),
                                        phiRef=rng.uniform(0, 6), f_ref=20.0, sky_average=1))
    return f, psd, srcs


def normalised(F, R):
    d = np.sqrt(np.abs(np.einsum("sii->si", R)))
    return np.abs(F - R) / (d[:, :, None] * d[:, None, :])


@pytest.mark.parametrize("order", [2, 4])
def test_sky_averaged_fisher_math_vs_reference(oracle, order):
    path = os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so")
    hh = C.CDLL(path)
    f, psd, srcs = setup_case()
    ref = oracle.fisher_numerical_batch("IMRPhenomD", srcs, ["Hanford"], f, psd[None, :], 7, order=order, detector_index=0)
    got = np.zeros_like(ref)
    for i, s in enumerate(srcs):
        assert hh.hh_fisher_numerical(b"IMRPhenomD", b"Hanford", b"Hanford", 7, order, C.byref(s), f.ctypes.data_as(_dp), f.size,
                                      psd.ctypes.data_as(_dp), got[i].ctypes.data_as(_dp)) == 0
    assert np.all(np.isfinite(ref)) and np.all(np.diagonal(ref, axis1=1, axis2=2) > 0)
    assert normalised(got, ref).max() <= NORM_TOL
    # F_00 = (h|h): the SNR^2 of the sky-averaged amplitude; F_01 = F_02 = 0 (amplitude and phase parameters decouple)
    assert np.abs(got[:, 0, 1]).max() <= 1e-6 * got[:, 0, 0].max() and np.abs(got[:, 0, 2]).max() <= 1e-6 * np.sqrt(got[:, 0, 0] * got[:, 2, 2]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 4])
def test_sky_averaged_fisher_vs_reference(ctx, oracle, order):
    f, psd, srcs = setup_case()
    ctx.set_network(["Hanford", "Livingston"], f, np.stack([psd, 2.0 * psd]))
    ref = oracle.fisher_numerical_batch("IMRPhenomD", srcs, ["Hanford"], f, psd[None, :], 7, order=order, detector_index=0)
    got = ctx.fisher_numerical_batch("IMRPhenomD", srcs, 7, order=order, detector_index=0)
    assert np.all(np.isfinite(got))
    assert normalised(got, ref).max() <= NORM_TOL_GPU
    assert np.median(normalised(got, ref)) <= 1e-6
    # the second detector's PSD is twice the first's: the matrix halves (to the rounding of the division)
    got2 = ctx.fisher_numerical_batch("IMRPhenomD", srcs, 7, order=order, detector_index=1)
    assert np.allclose(got2, 0.5 * got, rtol=1e-12, atol=0)
    # symmetric, and the same matrices whatever the batch they are computed in
    assert np.array_equal(got, np.swapaxes(got, 1, 2))
    assert np.array_equal(ctx.fisher_numerical_batch("IMRPhenomD", srcs[2:3], 7, order=order, detector_index=0)[0], got[2])


@pytest.mark.gpu
def test_sky_averaged_fisher_argument_errors(ctx):
    f, psd, srcs = setup_case()
    ctx.set_network(["Hanford"], f, psd[None, :])
    pointed = abi.source_defaults(mass1=30., mass2=20., Luminosity_Distance=400., f_ref=20.0)
    for kwargs, code in ((dict(method="IMRPhenomD", sources=srcs + [pointed], dimension=7, detector_index=0), abi.ERR_ARG),
                         (dict(method="IMRPhenomD", sources=srcs, dimension=11, detector_index=0), abi.ERR_ARG),
                         (dict(method="IMRPhenomPv2", sources=srcs, dimension=7, detector_index=0), abi.ERR_UNSUPPORTED),
                         (dict(method="IMRPhenomD", sources=srcs, dimension=7, detector_index=-1), abi.ERR_ARG)):
        with pytest.raises(engine.GwatB200Error) as e:
            ctx.fisher_numerical_batch(kwargs["method"], kwargs["sources"], kwargs["dimension"], order=4, detector_index=kwargs["detector_index"])
        assert e.value.code == code


# ---- the modified families behind the seven parameters (ppE, the theories mapped onto ppE, gIMR) --------------
def mod_cases():
    f, psd, srcs = setup_case()
    out = []
    for method, kw in (("ppE_IMRPhenomD_Inspiral", dict(Nmod=2, bppe=[-3.0, 1.0], betappe=[0.02, 0.3])),  # (beta_2 > 1/4 - eps: the one-sided rule)
                       ("ppE_IMRPhenomD_IMR", dict(Nmod=1, bppe=[-1.0], betappe=[0.01])),
                       ("dCS_IMRPhenomD", dict(Nmod=1, bppe=[-1.0], betappe=[1e-20])),
                       ("gIMRPhenomD", dict(Nmod_phi=1, phii=[4], delta_phi=[0.05], Nmod_beta=1, betai=[2], delta_beta=[0.02]))):
        ss = []
        for s in srcs[:3]:
            t = abi.Source()
            C.memmove(C.addressof(t), C.addressof(s), C.sizeof(s))
            for k, v in kw.items():
                if isinstance(v, list):
                    for i, x in enumerate(v):
                        getattr(t, k)[i] = x
                else:
                    setattr(t, k, v)
            ss.append(t)
        mods = kw.get("Nmod", 0) + kw.get("Nmod_phi", 0) + kw.get("Nmod_beta", 0)
        out.append((method, 7 + mods, ss))
    return f, psd, out


@pytest.mark.parametrize("order", [2, 4])
def test_sky_averaged_modified_families_math_vs_reference(oracle, order):
    hh = C.CDLL(os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so"))
    f, psd, cs = mod_cases()
    for method, dim, srcs in cs:
        ref = oracle.fisher_numerical_batch(method, srcs, ["Hanford"], f, psd[None, :], dim, order=order, detector_index=0)
        got = np.zeros_like(ref)
        for i, s in enumerate(srcs):
            assert hh.hh_fisher_numerical(method.encode(), b"Hanford", b"Hanford", dim, order, C.byref(s), f.ctypes.data_as(_dp), f.size,
                                          psd.ctypes.data_as(_dp), got[i].ctypes.data_as(_dp)) == 0, method
        if method == "dCS_IMRPhenomD":
            # ln A0 +- 1e-8 makes the luminosity distance negative (A0 ~ 1e-21), and the theory mapping takes its redshift: the reference's
            # own matrix is NaN in that row and column.  Same pattern here; the finite block agrees.
            assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(ref[:, 0, :]).all() and np.isfinite(ref[:, 1:, 1:]).all()
            ref, got = ref[:, 1:, 1:], got[:, 1:, 1:]
        assert np.all(np.isfinite(ref)), method
        # yardstick: the reference against itself on inputs moved by parts in 1e14 (tests/fisher_noise.py)
        floor = fisher_noise.reference_self_difference(oracle, method, srcs, ["Hanford"], f, psd[None, :], dim, order, detector_index=0)
        err = normalised(got, ref).reshape(len(srcs), -1).max(axis=1)
        assert np.all(err <= np.maximum(NORM_TOL, fisher_noise.FACTOR * floor)), (method, err, floor)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 4])
def test_sky_averaged_modified_families_vs_reference(ctx, oracle, order):
    f, psd, cs = mod_cases()
    ctx.set_network(["Hanford"], f, psd[None, :])
    for method, dim, srcs in cs:
        ref = oracle.fisher_numerical_batch(method, srcs, ["Hanford"], f, psd[None, :], dim, order=order, detector_index=0)
        got = ctx.fisher_numerical_batch(method, srcs, dim, order=order, detector_index=0)
        assert got.shape == ref.shape, method
        if method == "dCS_IMRPhenomD":
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            ref, got = ref[:, 1:, 1:], got[:, 1:, 1:]
        assert np.all(np.isfinite(got)), method
        floor = fisher_noise.reference_self_difference(oracle, method, srcs, ["Hanford"], f, psd[None, :], dim, order, detector_index=0)
        err = normalised(got, ref).reshape(len(srcs), -1).max(axis=1)
        assert np.all(err <= np.maximum(NORM_TOL_GPU, fisher_noise.FACTOR * floor)), (method, err, floor)
        assert np.median(normalised(got, ref)) <= 1e-5, method


@pytest.mark.gpu
def test_sky_averaged_nrt_is_refused_with_the_reason(ctx):
    f, psd, srcs = setup_case()
    ctx.set_network(["Hanford"], f, psd[None, :])
    with pytest.raises(engine.GwatB200Error) as e:
        ctx.fisher_numerical_batch("IMRPhenomD_NRT", srcs, 8, order=4, detector_index=0)
    assert e.value.code == abi.ERR_UNSUPPORTED and "ln(eta)" in str(e.value)
