"""Theory -> ppE mappings beyond dCS/EdGB (assign_mapping, src/ppE_utilities.cpp:158-359): EdGB_HO (== EdGB, a quirk of the
reference's if-chain), EdGB_HO_LO, EdGB_GHOv1-3, ExtraDimension, BHEvaporation, TVG, DipRad, NonComm, ModDispersion, PNSeries_ppE, ppEAlt.
Golden values from the reference's own code (tests/golden/make_golden.py --only-theories); 1e-10 on waveforms, 1e-9 on logL.
"""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from gw_analysis_tools_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
_dp = C.POINTER(C.c_double)
IDS = [c[0] for c in cases.THEORY_CASES]


def _p(a):
    return a.ctypes.data_as(_dp)


def _relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "theories_v1.npz"))


@pytest.fixture(scope="module")
def hh():
    path = os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so")
    if not os.path.exists(path):
        pytest.fail("tests/_build/libgwat_host_harness.so missing: run __graft_entry__.build()")
    return C.CDLL(path)


@pytest.mark.parametrize("case", cases.THEORY_CASES, ids=IDS)
def test_host_math_vs_golden(hh, gold, case):
    name, method, kw, gspec = case
    f = cases.grid(gspec)
    src = cases.source_from_bytes(gold[name + "/src"])
    o = [np.zeros(f.size) for _ in range(4)]
    assert hh.hh_fourier_waveform(method.encode(), C.byref(src), _p(f), f.size, *[_p(x) for x in o]) == 0
    assert _relerr(o[0] + 1j * o[1], gold[name + "/hp"]) <= 1e-10
    assert _relerr(o[2] + 1j * o[3], gold[name + "/hc"]) <= 1e-10


def test_edgb_ho_is_plain_edgb(gold):
    """The reference's mapping for "EdGB_HO_<model>" ends in the plain EdGB branch; the golden values show it."""
    wf = np.load(os.path.join(GOLD, "waveforms_v1.npz"))
    assert np.array_equal(gold["EdGB_HO/hp"], wf["EdGB/hp"])
    assert not np.array_equal(gold["EdGB_HO_LO/hp"], wf["EdGB/hp"])


def test_oracle_reproduces_theory_goldens(oracle, gold):
    for name, method, kw, gspec in cases.THEORY_CASES[::3]:
        f = cases.grid(gspec)
        src = cases.source_from_bytes(gold[name + "/src"])
        hp, hc = oracle.fourier_waveform(method, src, f)
        assert np.array_equal(hp, gold[name + "/hp"]) and np.array_equal(hc, gold[name + "/hc"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases.THEORY_CASES, ids=IDS)
def test_cuda_vs_golden(ctx, gold, case):
    name, method, kw, gspec = case
    wf = np.load(os.path.join(GOLD, "waveforms_v1.npz"))
    f = cases.grid(gspec)
    src = cases.source_from_bytes(gold[name + "/src"])
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    ctx.set_network(cases.DETECTORS, f, psd, cases.maximized_data(wf, gspec))
    hp, hc = ctx.fourier_waveform_batch(method, [src])
    assert _relerr(hp[0], gold[name + "/hp"]) <= 1e-10
    assert _relerr(hc[0], gold[name + "/hc"]) <= 1e-10
    ll = ctx.loglike_batch(method, [src])[0]
    ref = float(gold[name + "/logL"])
    assert abs(ll - ref) <= 1e-9 * abs(ref), (ll, ref)


@pytest.mark.gpu
def test_theory_mcmc_vectors_and_fisher(ctx, oracle):
    """Sampling-vector entry point for a two-parameter theory (unit conversion of sqrt(alpha), generic second parameter) and
    its Fisher matrix, against the compiled reference."""
    from gw_analysis_tools_b200 import abi
    from gw_analysis_tools_b200 import sampler as smp
    wl = workloads.make(1, W=8, L=2048)
    method = "EdGB_GHOv2_IMRPhenomD"
    mod = abi.mod_defaults(ppE_Nmod=2, bppe=[-7., -5.])
    rng = np.random.default_rng(4)
    params = np.concatenate([wl.params, rng.uniform(1., 5., (8, 1)), rng.uniform(0.5, 2., (8, 1))], axis=1)
    inj = np.concatenate([wl.inj, [2.0, 1.0]])
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(method, inj[None, :], wl.gmst, mod)
    src[0].tc = wl.T_segment - src[0].tc
    data = ctx.coherent_response_batch(method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    got = ctx.loglike_mcmc_batch(method, params, wl.gmst, wl.T_segment, mod)
    want = oracle.loglike_mcmc_batch(method, mod, params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, data)
    assert (np.abs(got - want) / np.abs(want)).max() <= 1e-9
    F, vals, vecs = smp.mcmc_fisher_batch(ctx, method, params[:2], wl.gmst, order=4, mod=mod)
    _, srcs = oracle.loglike_mcmc_batch(method, mod, params[:2], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, None, return_sources=True)
    for s in (srcs[0], srcs[1]):
        s.tc = wl.T_segment - s.tc
    ref = oracle.fisher_numerical_batch("MCMC_" + method, srcs, wl.detectors, wl.f, wl.psd, 13, 4, detector_index=-1, reference_index=0)
    from oracle import ptmcmc_ref
    # rows/columns of the coupling itself are left out: the stencil steps alpha^2 ~ 1e-20 s^4 by eps = 1e-8 (absolute, as the
    # reference does, src/fisher.cpp:340-543), which overflows on both sides and says nothing about either
    # ... and so is the row of the generic second parameter: a derivative of ~1e-20 relative size on top of eps = 1e-8
    # cancellation noise, different on every machine
    keep = np.arange(11)
    for i in range(2):
        want_f = ptmcmc_ref.fisher_transformations(ref[i], False, True, 2, params[i])[np.ix_(keep, keep)]
        got_f = F[i][np.ix_(keep, keep)]
        dg = np.sqrt(np.abs(np.diag(want_f)))
        nerr = np.abs(got_f - want_f) / np.outer(dg, dg)
        assert np.median(nerr) <= 1e-6 and nerr.max() <= 1e-4, (np.median(nerr), nerr.max())


# ---- cosmologies of Z_from_DL other than the default (include/gwat/D_Z_Config.h: cosmos[]) ----------------------------------
COSMO_METHODS = ["dCS_IMRPhenomD", "EdGB_IMRPhenomD", "ModDispersion_IMRPhenomD", "ExtraDimension_IMRPhenomD"]


def _cosmo_source(gold, method, index):
    name = [c[0] for c in (cases.THEORY_CASES + cases.CASES) if c[1] == method][0]
    g = gold if name + "/src" in gold else np.load(os.path.join(GOLD, "waveforms_v1.npz"))
    src = cases.source_from_bytes(g[name + "/src"])
    src.cosmology = index
    src.Luminosity_Distance = 2500.0  # z ~ 0.4-0.45: the cosmologies differ in the third digit of z
    return src, [c for c in (cases.THEORY_CASES + cases.CASES) if c[0] == name][0][3]


@pytest.mark.parametrize("index", [1, 2, 3, 4, 5], ids=["PLANCK13", "WMAP9", "WMAP7", "WMAP5", "TESTING_COSMOLOGY"])
def test_cosmologies_host_math_vs_reference(hh, oracle, gold, index):
    for method in COSMO_METHODS:
        src, gspec = _cosmo_source(gold, method, index)
        f = cases.grid(gspec)
        o = [np.zeros(f.size) for _ in range(4)]
        assert hh.hh_fourier_waveform(method.encode(), C.byref(src), _p(f), f.size, *[_p(x) for x in o]) == 0
        hp, hc = oracle.fourier_waveform(method, src, f)
        assert _relerr(o[0] + 1j * o[1], hp) <= 1e-10, method
        src0, _ = _cosmo_source(gold, method, 0)
        hp0, _ = oracle.fourier_waveform(method, src0, f)
        if index != 5:  # (TESTING_COSMOLOGY carries PLANCK15's numbers)
            assert _relerr(hp0, hp) > 1e-6, method  # the cosmology is really read


def test_cosmology_names():
    from gw_analysis_tools_b200 import engine
    lib = engine.load_library()
    for i, n in enumerate(["PLANCK15", "PLANCK13", "WMAP9", "WMAP7", "WMAP5", "TESTING_COSMOLOGY"]):
        assert lib.gwat_b200_cosmology_index(n.encode()) == i
        assert lib.gwat_b200_cosmology_index(n.lower().encode()) == i  # Z_from_DL upper-cases the name
    assert lib.gwat_b200_cosmology_index(b"EINSTEIN_DE_SITTER") == -1


@pytest.mark.gpu
@pytest.mark.parametrize("index", [1, 2, 4], ids=["PLANCK13", "WMAP9", "WMAP5"])
def test_cosmologies_cuda_vs_reference(ctx, oracle, gold, index):
    for method in COSMO_METHODS:
        src, gspec = _cosmo_source(gold, method, index)
        f = cases.grid(gspec)
        psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
        ctx.set_network(cases.DETECTORS, f, psd)
        hp, hc = ctx.fourier_waveform_batch(method, [src])
        rp, rc = oracle.fourier_waveform(method, src, f)
        assert _relerr(hp[0], rp) <= 1e-10, method
        assert _relerr(hc[0], rc) <= 1e-10, method
