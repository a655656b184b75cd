"""libgwat_b200_dropin.so: the reference's own C++ link-level symbols (T = double) on top of the C ABI.

tests/cxx/dropin_caller.cpp is a program written the way GWAT's own programs are (GWAT headers, GWAT types, GWAT calls).
__graft_entry__.build() compiles it ONCE against the reference's headers and links it twice:
  tests/_build/dropin_caller_ref   against oracle/_ref/libgwat_ref.so alone (the reference's CPU code runs every call)
  tests/_build/dropin_caller_b200  with libgwat_b200_dropin.so ahead of it on the link line (the hot-path symbols resolve to
                                   the GPU library -- also for calls made from inside libgwat_ref.so -- everything else to the reference)
CPU tier: the mangled names this library defines are exactly names the reference library defines.
GPU tier: both programs print the same numbers.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "gw_analysis_tools_b200", "libgwat_b200_dropin.so")
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libgwat_ref.so")
BUILD = os.path.join(ROOT, "tests", "_build")

HOT_PATH = ["fourier_waveform<double>(", "fourier_waveform(double*, int, std::complex<double>*", "fourier_waveform(double*, int, double*, double*, std::",
            "fourier_waveform(double*, int, double*, double*, double*, double*, double*", "fourier_detector_response<double>(",
            "create_coherent_GW_detection<double>(", "create_coherent_GW_detection_reuse_WF<double>(", "Log_Likelihood_internal(",
            "MCMC_likelihood_extrinsic(", "MCMC_likelihood_wrapper(", "MCMC_fisher_wrapper(", "fisher_numerical(",
            "calculate_snr(std::"]


def _defined(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    return {ln.split()[2] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] in "TW"}


def _need(path):
    if not os.path.exists(path):
        pytest.skip(os.path.relpath(path, ROOT) + " not built (needs the reference headers: __graft_entry__.build() in the build container)")


def test_dropin_defines_the_reference_link_symbols():
    _need(DROPIN)
    _need(REFLIB)
    mine, ref = _defined(DROPIN), _defined(REFLIB)
    shared = sorted(mine & ref)
    demangled = subprocess.run(["c++filt"], input="\n".join(shared), check=True, capture_output=True, text=True).stdout.splitlines()
    for want in HOT_PATH:
        assert any(want in d for d in demangled), "not defined with the reference's mangled name: " + want
    # and nothing else of the reference is shadowed by accident (std:: template instances are weak and harmless)
    extra = [d for d in demangled if not any(w in d for w in HOT_PATH) and not d.startswith("std::")]
    assert extra == [], extra


def _run(name):
    exe = os.path.join(BUILD, name)
    _need(exe)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=600, env=env).stdout
    vals = {}
    for ln in out.splitlines():
        k, v = ln.rsplit(" ", 1)
        vals[k] = float(v)
    return vals


def test_reference_link_runs_on_cpu():
    ref = _run("dropin_caller_ref")
    assert ref["status_wf_IMRPhenomD"] == 1 and np.isfinite(ref["MCMC_likelihood_extrinsic_Pv2"])
    # the legacy overloads do nothing for a precessing model
    assert ref["status_legacy_IMRPhenomPv2"] == 1 and ref["legacy_complex_re_IMRPhenomPv2[100]"] == 0.0


@pytest.mark.gpu
def test_same_program_same_numbers_through_the_gpu_library():
    ref, got = _run("dropin_caller_ref"), _run("dropin_caller_b200")
    only_b200 = {k for k in got if k.startswith("MCMC_likelihood_wrapper") or k.startswith("MCMC_fisher_wrapper")}
    assert set(ref) == set(got) - only_b200
    groups = {}
    for k in ref:
        groups.setdefault(k.split("[")[0], []).append(k)
    for g, keys in groups.items():
        a, b = np.array([ref[k] for k in keys]), np.array([got[k] for k in keys])
        scale = np.abs(a).max()
        if g.startswith("status"):
            assert np.array_equal(a, b), g
        elif g.startswith(("hp_", "hc_", "legacy_", "single_", "coherent_")):
            # strain samples: 1e-10 of the largest sample of the group x 100 (the samples span two decades below the peak of |h|)
            assert np.abs(a - b).max() <= 1e-8 * scale, (g, a, b)
        elif "fisher" in g and "diag" in g and "offdiag" not in g:
            assert (np.abs(a - b) / np.abs(a)).max() <= 2e-5, (g, a, b)  # the reference's own FMA-vs-non-FMA floor for these stencils
        elif "fisher" in g:
            pass  # off-diagonals: normalised below
        else:
            assert np.abs(a - b).max() <= 1e-9 * scale, (g, a, b)  # likelihoods
    for off, i, j, diag in (("fisher_sum_offdiag_7_8", 7, 8, "fisher_sum_diag"), ("fisher_numerical_o2_offdiag_6_7", 6, 7, "fisher_numerical_o2_diag")):
        norm = np.sqrt(ref["%s[%d]" % (diag, i)] * ref["%s[%d]" % (diag, j)])
        assert abs(ref[off] - got[off]) / norm <= 2e-5, off
    # the samplers' callbacks: the wrapper = the chain of reference calls it stands for
    assert abs(got["MCMC_likelihood_wrapper"] - ref["callback_chain"]) <= 1e-9 * abs(ref["callback_chain"])
    for i in range(11):
        a, b = ref["fisher_sum_diag[%d]" % i], got["MCMC_fisher_wrapper_diag[%d]" % i]
        assert abs(a - b) <= 2e-5 * abs(a), (i, a, b)
    # ... on a Fisher grid of its own (user_param->fisher_freq, fisher_PSD, fisher_length; here geometric, 384 bins), and back on the data grid
    for i in range(11):
        a, b = ref["fisher_owngrid_sum_diag[%d]" % i], got["MCMC_fisher_wrapper_owngrid_diag[%d]" % i]
        assert abs(a - b) <= 2e-5 * abs(a), (i, a, b)
    assert abs(ref["fisher_owngrid_sum_diag[7]"] - ref["fisher_sum_diag[7]"]) > 1e-3 * abs(ref["fisher_sum_diag[7]"])  # (another grid, another matrix)
    assert got["MCMC_likelihood_wrapper_after_owngrid"] == got["MCMC_likelihood_wrapper"]
    # ... and in an intrinsic run: the tc/phic-maximised likelihood and the sky-averaged Fisher of the 4-parameter set
    assert abs(got["MCMC_likelihood_wrapper_intrinsic"] - ref["intrinsic_callback_chain"]) <= 1e-9 * abs(ref["intrinsic_callback_chain"])
    for i in range(4):
        a, b = ref["intrinsic_fisher_sum_diag[%d]" % i], got["MCMC_fisher_wrapper_intrinsic_diag[%d]" % i]
        assert abs(a - b) <= 2e-5 * abs(a), (i, a, b)
    assert ref["intrinsic_fisher_sum_diag[1]"] == 4.0 and got["MCMC_fisher_wrapper_intrinsic_diag[2]"] == 0.25  # the prior terms replace these
    # (entry (1,1) is a prior term now, so the ln Mc - eta entry is compared with itself: the two are 90 % correlated)
    assert abs(ref["intrinsic_fisher_sum_offdiag_0_1"] - got["MCMC_fisher_wrapper_intrinsic_offdiag_0_1"]) <= 2e-5 * abs(ref["intrinsic_fisher_sum_offdiag_0_1"])
