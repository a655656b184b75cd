"""CPU tier: the kernels' GWAT_HD mathematics, compiled as plain C++ (tests/host_harness.cpp), against the golden vectors
generated from the reference's own code, and the oracle itself against the same vectors.

Tolerances are BASELINE.json's: waveform <= 1e-10 of max|h|, logL <= 1e-9 relative.  Fisher matrices are compared with the
normalised measure max_ij |dF_ij| / sqrt(F_ii F_jj).  Finite differences with eps = 1e-8 amplify rounding noise by 1e8, so
the reference does not reproduce ITSELF to 1e-6: tests/golden/fisher_noise_v2.npz stores, for every golden matrix, the
reference's self-difference (its FMA-contracted build and four re-evaluations on inputs moved by parts in 1e14,
tests/fisher_noise.py).  The bounds used here: median_ij <= 1e-6 and max_ij <= max(1e-6, 3 x that self-difference).
"""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from gw_analysis_tools_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
WF_TOL = 1e-10
FISHER_NORM_TOL = 1e-6
import fisher_noise  # noqa: E402

FISHER_NOISE_FACTOR = fisher_noise.FACTOR
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hh():
    path = os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build_harness()
    return C.CDLL(path)


@pytest.fixture(scope="module")
def gold_wf():
    return np.load(os.path.join(GOLD, "waveforms_v1.npz"))


@pytest.fixture(scope="module")
def gold_fisher():
    return np.load(os.path.join(GOLD, "fisher_v1.npz"))


def _p(a):
    return a.ctypes.data_as(_dp)


def _relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_waveform_and_response_vs_golden(hh, gold_wf, case):
    name, method, kw, gspec = case
    f = cases.grid(gspec)
    src = cases.source_from_bytes(gold_wf[name + "/src"])
    L = f.size
    o = [np.zeros(L) for _ in range(4)]
    assert hh.hh_fourier_waveform(method.encode(), C.byref(src), _p(f), L, *[_p(x) for x in o]) == 0
    assert _relerr(o[0] + 1j * o[1], gold_wf[name + "/hp"]) <= WF_TOL
    assert _relerr(o[2] + 1j * o[3], gold_wf[name + "/hc"]) <= WF_TOL
    D = len(cases.DETECTORS)
    re, im = np.zeros((D, L)), np.zeros((D, L))
    dets = (C.c_char_p * D)(*[d.encode() for d in cases.DETECTORS])
    assert hh.hh_coherent_response(method.encode(), C.byref(src), D, dets, _p(f), L, 1, _p(re), _p(im)) == 0
    for d in range(D):
        assert _relerr(re[d] + 1j * im[d], gold_wf[name + "/resp"][d]) <= WF_TOL
    if name + "/single_L" in gold_wf:
        assert hh.hh_coherent_response(method.encode(), C.byref(src), D, dets, _p(f), L, 0, _p(re), _p(im)) == 0
        assert _relerr(re[1] + 1j * im[1], gold_wf[name + "/single_L"]) <= WF_TOL


@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_cooperative_setup_dataflow_is_bit_identical(hh, gold_wf, case):
    """k_setup's roles and steps (gwat_setup_coop.h) run on the host, role by role on separate carry structures with everything
    a role does not own poisoned: the merged coefficient record equals walker_setup's bit for bit."""
    name, method, kw, gspec = case
    src = cases.source_from_bytes(gold_wf[name + "/src"])
    layout = (C.c_int * 6)()
    hh.hh_walkercoef_layout(layout)
    o_d, o_p, o_fac, o_det, o_valid, det_words = list(layout)
    dets = ["Hanford", "Livingston", "Virgo"]
    arr = (C.c_char_p * 3)(*[d.encode() for d in dets])
    a, b = np.zeros(512, dtype=np.uint64), np.zeros(512, dtype=np.uint64)
    n = hh.hh_setup_coop(method.encode(), C.byref(src), 3, arr, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert n > 0, n
    a, b = a[:n], b[:n]
    poison = np.uint64(0xffffffffffffffff)
    # words the sequential flow defines must be defined identically by the cooperative one; detectors beyond D stay untouched
    defined = b != poison
    defined[o_det + 3 * det_words:o_valid] = False
    assert defined[o_d:o_d + 20].all()  # (the poison pattern is not a value any field takes)
    # (an int field shares its word with 4 bytes of padding, which stay poisoned in the sequential flow: compare the int alone)
    half = np.uint64(0xffffffff)
    int_words = defined & ((b >> np.uint64(32)) == half)
    full = defined & ~int_words
    assert np.array_equal(a[full], b[full]), np.nonzero(full & (a != b))[0]
    assert np.array_equal(a[int_words] & half, b[int_words] & half), np.nonzero(int_words & ((a & half) != (b & half)))[0]
    # and the cooperative flow leaves nothing undefined that the per-bin code of this family reads
    assert not (a[full] == poison).any()


def test_modified_families_differ_from_gr(gold_wf):
    """The modification terms are really exercised: each modified case is far (>> tolerance) from its GR counterpart."""
    for name, base in [("ppE_ins", "D_bbh"), ("ppE_imr", "D_bbh"), ("gIMR", "D_bbh"), ("gIMR_log", "D_bbh"), ("dCS", "D_bbh"),
                       ("EdGB", "D_low"), ("ppE_P_ins", "P_full"), ("gIMR_P", "P_full"), ("dCS_P", "P_full")]:
        assert _relerr(gold_wf[name + "/hp"], gold_wf[base + "/hp"]) > 1e-4, name


@pytest.mark.parametrize("case", cases.FISHER_CASES, ids=[c[0] for c in cases.FISHER_CASES])
def test_fisher_vs_golden(hh, gold_fisher, case):
    gold_noise = np.load(os.path.join(GOLD, "fisher_noise_v2.npz"))
    name, method, kw, dim = case
    f = cases.grid(cases.FISHER_GRID)
    psd = workloads.aligo_analytic_psd(f)
    src = cases.source_from_bytes(gold_fisher[name + "/src"])
    for order in (2, 4):
        for det in cases.DETECTORS[:2]:
            out = np.zeros((dim, dim))
            rc = hh.hh_fisher_numerical(method.encode(), det.encode(), b"Hanford", dim, order, C.byref(src), _p(f), f.size,
                                        _p(psd), _p(out))
            assert rc == 0
            ref = gold_fisher["%s/o%d/%s" % (name, order, det)]
            dg = np.sqrt(np.abs(np.diag(ref)))
            worst_floor = float(gold_noise["%s/o%d/%s" % (name, order, det)])
            nerr = np.abs(out - ref) / np.outer(dg, dg)
            assert np.median(nerr) <= FISHER_NORM_TOL, (name, order, det, np.median(nerr))
            assert nerr.max() <= max(FISHER_NORM_TOL, FISHER_NOISE_FACTOR * worst_floor), (name, order, det, nerr.max())
            assert np.allclose(out, out.T)


def test_unknown_method_is_rejected(hh):
    f = cases.grid(cases.GRID_BBH)
    src = cases.source(cases.BBH)
    o = [np.zeros(f.size) for _ in range(4)]
    for bad in ("IMRPhenomXYZ", "", "IMRPhenomPv3", "EA_IMRPhenomD_NRT"):
        assert hh.hh_fourier_waveform(bad.encode(), C.byref(src), _p(f), f.size, *[_p(x) for x in o]) == -2


LL_TOL = 1e-9  # BASELINE.json: log-likelihood relative error


@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_fused_likelihood_vs_golden(hh, gold_wf, case):
    """gwat_like.h -- the fused per-bin likelihood algebra and the fixed summation tree of k_loglike (units of bins, 256
    threads per unit) -- compiled as C++, against the reference's stored logL; and the value must not depend on the unit size
    beyond rounding."""
    name, method, kw, gspec = case
    f = cases.grid(gspec)
    src = cases.source_from_bytes(gold_wf[name + "/src"])
    psd = np.ascontiguousarray(np.tile(workloads.aligo_analytic_psd(f), (3, 1)))
    data = cases.derived_data(gold_wf[name + "/resp"])
    dre, dim = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    dets = (C.c_char_p * 3)(*[d.encode() for d in cases.DETECTORS])
    ref = float(gold_wf[name + "/logL"])
    got = []
    for unit in (0, 256, 4096):
        out = np.zeros(1)
        rc = hh.hh_loglike(method.encode(), 1, C.byref(src), 3, dets, _p(f), f.size, _p(psd), _p(dre), _p(dim), None, 0, 0, unit, _p(out))
        assert rc == 0
        assert abs(out[0] - ref) <= LL_TOL * abs(ref), (unit, out[0], ref)
        got.append(out[0])
    assert max(got) - min(got) <= 1e-12 * abs(ref)


def test_grid_tables_reproduce_glibc_pow():
    """f^(fl(1/6)) * M^(fl(1/6)) in double-double, rounded once, equals glibc's pow(M*f, 1./6.) to <= 1 ulp."""
    rng = np.random.default_rng(3)
    f = rng.uniform(5, 4096, 20000)
    M = rng.uniform(2, 200, 20000) * 4.925491025543576e-06
    ld = np.longdouble
    sf = np.power(f.astype(ld), ld(1.0 / 6.0))
    sm = np.power(M.astype(ld), ld(1.0 / 6.0))
    mine = (sf * sm).astype(np.float64)
    ref = np.power(M * f, 1.0 / 6.0)
    ulp = np.abs(mine - ref) / np.spacing(ref)
    assert ulp.max() <= 1.0
    assert (ulp == 0).mean() > 0.85


# ---- the oracle itself is pinned by the same vectors ---------------------------------------------------------------------

def test_oracle_reproduces_golden(oracle, gold_wf):
    for name, method, kw, gspec in cases.CASES[::4]:
        f = cases.grid(gspec)
        src = cases.source_from_bytes(gold_wf[name + "/src"])
        hp, hc = oracle.fourier_waveform(method, src, f)
        assert np.array_equal(hp, gold_wf[name + "/hp"]) and np.array_equal(hc, gold_wf[name + "/hc"])
        resp = oracle.coherent_response(method, src, cases.DETECTORS, f)
        psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
        ll = oracle.loglike_batch(method, [src], cases.DETECTORS, f, psd, cases.derived_data(resp))[0]
        assert ll == float(gold_wf[name + "/logL"])


def test_oracle_sample_value_from_survey(oracle):
    """SURVEY.md Appendix A quotes logL = 1.495080331741781e+04 for this configuration of the reference."""
    from gw_analysis_tools_b200 import abi
    f = 20 + 0.25 * np.arange(8192)
    psd = np.tile(oracle.populate_noise(f, "aLIGO_analytic") ** 2, (3, 1))
    inj = abi.source_defaults(**cases.BBH)
    data = oracle.coherent_response("IMRPhenomD", inj, cases.DETECTORS, f)
    tmpl = abi.source_defaults(**dict(cases.BBH, mass1=36.4 * 1.0005))
    ll = oracle.loglike_batch("IMRPhenomD", [tmpl], cases.DETECTORS, f, psd, data)[0]
    assert abs(ll - 1.495080331741781e+04) <= 1e-9 * 1.5e4
    assert np.allclose(oracle.populate_noise(f, "aLIGO_analytic") ** 2, workloads.aligo_analytic_psd(f), rtol=1e-15, atol=0)
