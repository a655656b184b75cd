"""include/gwat_b200_cxx.hpp -- the C++ forwarding layer with the reference's own function names and argument lists.

CPU tier: the header compiles against the small stand-in struct and, when the reference tree is present (build container),
against the real gen_params_base<double> of include/gwat/util.h.
GPU tier: tests/cxx/adapter_shim.cpp calls it with reference-style arguments; results against the golden vectors.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from gw_analysis_tools_b200 import abi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SHIM = os.path.join(ROOT, "tests", "cxx", "adapter_shim.cpp")
INCS = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "cxx")]
REF = "/root/reference"


def test_header_compiles_against_stand_in():
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror"] + INCS + [SHIM], check=True)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "gwat")), reason="reference headers not on this box")
def test_header_compiles_against_reference_gen_params():
    stubs = os.path.join(ROOT, "standins")
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-DGWAT_CXX_REAL_HEADERS", "-I" + stubs, "-I" + os.path.join(stubs, "gwatcfg"),
                    "-I" + os.path.join(REF, "include", "gwat"), "-I" + os.path.join(REF, "include")] + INCS + [SHIM], check=True)


@pytest.fixture(scope="module")
def shim():
    path = os.path.join(ROOT, "tests", "_build", "libgwat_cxx_adapter.so")
    if not os.path.exists(path):
        pytest.fail("tests/_build/libgwat_cxx_adapter.so missing: run __graft_entry__.build()")
    lib = C.CDLL(path)
    lib.cxa_loglike.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


ADAPTER_CASES = [c for c in cases.CASES if c[0] in ("D_bbh", "P_full", "P_nrt", "ppE_imr", "gIMR_log", "gIMR_P", "dCS", "NRT_love", "ppE_NRT_ins")]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ADAPTER_CASES, ids=[c[0] for c in ADAPTER_CASES])
def test_reference_style_calls_vs_golden(shim, case):
    name, method, kw, gspec = case
    gold = np.load(os.path.join(GOLD, "waveforms_v1.npz"))
    f = cases.grid(gspec)
    L = len(f)
    src = cases.source_from_bytes(gold[name + "/src"])
    m = method.encode()
    wf = np.zeros((4, L))
    assert shim.cxa_fourier_waveform(C.byref(src), m, _p(f), L, _p(wf)) == 1
    assert _relerr(wf[0] + 1j * wf[1], gold[name + "/hp"]) <= 1e-10
    assert _relerr(wf[2] + 1j * wf[3], gold[name + "/hc"]) <= 1e-10
    re, im = np.zeros((3, L)), np.zeros((3, L))
    assert shim.cxa_coherent_response(C.byref(src), m, b"Hanford,Livingston,Virgo", _p(f), L, _p(re), _p(im)) == 1
    resp_gold = gold[name + "/resp"]
    for d in range(3):
        assert _relerr(re[d] + 1j * im[d], resp_gold[d]) <= 1e-10
    if name + "/single_L" in gold:
        one = np.zeros((2, L))
        assert shim.cxa_fourier_detector_response(C.byref(src), m, b"Livingston", _p(f), L, _p(one)) == 1
        assert _relerr(one[0] + 1j * one[1], gold[name + "/single_L"]) <= 1e-10
    # MCMC_likelihood_extrinsic flips tc -> T - tc; hand it the pre-image so that the golden logL applies
    T = 1.0 / (f[1] - f[0])
    trial = cases.source_from_bytes(gold[name + "/src"])
    trial.tc = T - src.tc
    data = cases.derived_data(resp_gold)
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    dre, dim_ = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    ll = shim.cxa_loglike(C.byref(trial), m, b"Hanford,Livingston,Virgo", _p(f), _p(psd), _p(dre), _p(dim_), L, C.c_double(-1.0))
    ref = float(gold[name + "/logL"])
    # T - (T - tc) is not exactly tc: allow the phase rounding of 2 pi f_max * ulp(T)
    assert abs(ll - ref) <= 1e-9 * abs(ref), (ll, ref)


@pytest.mark.gpu
def test_reference_style_fisher_vs_golden(shim):
    gold = np.load(os.path.join(GOLD, "fisher_v1.npz"))
    name, method, kw, dim = cases.FISHER_CASES[0]
    f = cases.grid(cases.FISHER_GRID)
    psd = workloads.aligo_analytic_psd(f)
    src = cases.source_from_bytes(gold[name + "/src"])
    out = np.zeros((dim, dim))
    assert shim.cxa_fisher_numerical(C.byref(src), method.encode(), b"Hanford", _p(f), _p(psd), len(f), dim, 4, _p(out)) == 1
    ref = gold["%s/o4/Hanford" % name]
    dg = np.sqrt(np.abs(np.diag(ref)))
    nerr = np.abs(out - ref) / np.outer(dg, dg)
    floor = float(gold["%s/o4/Hanford/noise" % name])
    assert np.median(nerr) <= 1e-6 and nerr.max() <= max(1e-6, 12 * floor)


@pytest.mark.gpu
def test_unknown_method_returns_failure_not_exit(shim):
    f = cases.grid(cases.GRID_BBH)
    src = cases.source(cases.BBH)
    wf = np.zeros((4, len(f)))
    assert shim.cxa_fourier_waveform(C.byref(src), b"IMRPhenomXYZ", _p(f), len(f), _p(wf)) == 0
