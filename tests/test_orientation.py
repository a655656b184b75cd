"""equatorial_orientation and horizon_coord: the one-detector entry points honour them as the reference's
fourier_detector_response / calculate_snr do (gwat_orient.h), the coherent response and the likelihoods read incl_angle, psi, RA,
DEC as given, as create_coherent_GW_detection_reuse_WF does."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from gw_analysis_tools_b200 import abi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def hh():
    return C.CDLL(os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so"))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "waveforms_v1.npz"))


def _sources(gold, name, n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        s = cases.source_from_bytes(gold[name + "/src"])
        s.equatorial_orientation = 1
        s.theta_l = rng.uniform(0.2, np.pi - 0.2)
        s.phi_l = rng.uniform(0, 2 * np.pi)
        s.RA = rng.uniform(0, 2 * np.pi)
        s.DEC = rng.uniform(-1.2, 1.2)
        s.incl_angle = 9.0  # garbage on purpose: the transform must replace both
        s.psi = 9.0
        if name.startswith("P_"):
            s.chip = rng.uniform(0.05, 0.8)  # PhenomPv2_JSF_from_params reads the reduced parameters
            s.phip = rng.uniform(0, 2 * np.pi)
        out.append(s)
    return out


@pytest.mark.parametrize("name,method", [("D_bbh", "IMRPhenomD"), ("P_full", "IMRPhenomPv2"), ("NRT_love", "IMRPhenomD_NRT")])
def test_transform_orientation_coords_vs_reference(hh, oracle, gold, name, method):
    for s in _sources(gold, name, 16, 3):
        incl_ref, psi_ref = oracle.transform_orientation_coords(method, "Hanford", s)
        t = abi.Source()
        C.memmove(C.addressof(t), C.addressof(s), C.sizeof(s))
        assert hh.hh_transform_orientation(method.encode(), C.byref(t)) == 0
        assert abs(t.incl_angle - incl_ref) <= 1e-13 * max(1.0, abs(incl_ref))
        assert abs(t.psi - psi_ref) <= 1e-12 * max(1.0, abs(psi_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("name,method", [("D_bbh", "IMRPhenomD"), ("P_full", "IMRPhenomPv2")])
def test_one_detector_response_with_equatorial_orientation(ctx, oracle, gold, name, method):
    gspec = [c for c in cases.CASES if c[0] == name][0][3]
    f = cases.grid(gspec)
    ctx.set_network(["Hanford", "Virgo"], f, np.tile(workloads.aligo_analytic_psd(f), (2, 1)))
    srcs = _sources(gold, name, 6, 11)
    got = ctx.fourier_detector_response_batch(method, "Virgo", srcs)
    for k, s in enumerate(srcs):
        ref = oracle.fourier_detector_response(method, "Virgo", s, f)
        assert np.abs(got[k] - ref).max() <= 1e-10 * np.abs(ref).max(), k
    # the coherent response does not look at the flag: garbage incl_angle / psi are used as given, as in the reference
    coh = ctx.coherent_response_batch(method, srcs[:2])
    for k in range(2):
        ref = oracle.coherent_response(method, srcs[k], ["Hanford", "Virgo"], f)
        assert np.abs(coh[k] - ref).max() <= 1e-10 * np.abs(ref).max()


@pytest.mark.gpu
def test_one_detector_response_in_horizon_coordinates(ctx, oracle, gold):
    name, method = "D_bbh", "IMRPhenomD"
    gspec = [c for c in cases.CASES if c[0] == name][0][3]
    f = cases.grid(gspec)
    ctx.set_network(["Livingston", "ET1"], f, np.tile(workloads.aligo_analytic_psd(f), (2, 1)))
    rng = np.random.default_rng(5)
    srcs = []
    for _ in range(4):
        s = cases.source_from_bytes(gold[name + "/src"])
        s.horizon_coord = 1
        s.theta, s.phi, s.psi = rng.uniform(0.1, 3.0), rng.uniform(0, 6.28), rng.uniform(0, 3.14)
        srcs.append(s)
    srcs.append(cases.source_from_bytes(gold[name + "/src"]))  # a batch may mix both conventions
    for det in ("Livingston", "ET1"):
        got = ctx.fourier_detector_response_batch(method, det, srcs)
        for k, s in enumerate(srcs):
            ref = oracle.fourier_detector_response(method, det, s, f)
            assert np.abs(got[k] - ref).max() <= 1e-10 * np.abs(ref).max(), (det, k)


@pytest.mark.gpu
def test_snr_with_equatorial_orientation(ctx, oracle, gold):
    name, method = "D_bbh", "IMRPhenomD"
    f = cases.grid([c for c in cases.CASES if c[0] == name][0][3])
    psd = oracle.populate_noise(f, "aLIGO_analytic") ** 2
    ctx.set_network(["Hanford"], f, psd[None, :])
    srcs = _sources(gold, name, 4, 21)
    got = ctx.snr_batch(method, srcs)
    for k, s in enumerate(srcs):
        ref = oracle.calculate_snr("aLIGO_analytic", "Hanford", method, s, f)
        assert abs(got[k] - ref) <= 1e-9 * ref, k


# ---- Fisher matrices with the direction of L as parameters (src/fisher.cpp:1851-1858, 2180-2187): theta_l and phi_l take the
# slots of psi and iota, and every stencil point derives incl_angle and psi again (src/waveform_util.cpp:947-949)
FISHER_EQ_CASES = [c for c in cases.FISHER_CASES if c[0] in ("F_D", "F_D_mcmc", "F_P_mcmc", "F_P_red")]


def _fisher_eq_source(name, seed):
    gold_fisher = np.load(os.path.join(GOLD, "fisher_v1.npz"))
    rng = np.random.default_rng(seed)
    s = cases.source_from_bytes(gold_fisher[name + "/src"])
    s.equatorial_orientation = 1
    s.theta_l = rng.uniform(0.3, 2.8)
    s.phi_l = rng.uniform(0, 2 * np.pi)
    s.incl_angle = 9.0  # garbage on purpose: neither may be read
    s.psi = 9.0
    return s


def _fisher_eq_check(got, oracle, method, src, f, psd2, dim, order, di):
    import fisher_noise
    ref = oracle.fisher_numerical_batch(method, [src], cases.DETECTORS[:2], f, psd2, dim, order=order, detector_index=di)[0]
    floor = fisher_noise.reference_self_difference(oracle, method, [src], cases.DETECTORS[:2], f, psd2, dim, order,
                                                   detector_index=di, runs=2)[0]
    nerr = fisher_noise.normalised_error(got, ref)
    assert np.all(np.isfinite(got))
    assert np.median(nerr) <= 1e-6, (method, order, di, np.median(nerr))
    assert nerr.max() <= max(1e-6, fisher_noise.FACTOR * floor), (method, order, di, nerr.max(), floor)
    # the orientation rows are really those of (theta_l, phi_l): the same source with the flag off gives another matrix
    plain = abi.Source()
    C.memmove(C.addressof(plain), C.addressof(src), C.sizeof(src))
    plain.equatorial_orientation = 0
    plain.incl_angle, plain.psi = 0.7, 0.4
    other = oracle.fisher_numerical_batch(method, [plain], cases.DETECTORS[:2], f, psd2, dim, order=order, detector_index=di)[0]
    assert fisher_noise.normalised_error(other, ref).max() > 1e-2


@pytest.mark.parametrize("case", FISHER_EQ_CASES, ids=[c[0] for c in FISHER_EQ_CASES])
def test_fisher_with_equatorial_orientation_host_math(hh, oracle, case):
    name, method, kw, dim = case
    f = cases.grid(cases.FISHER_GRID)
    psd = workloads.aligo_analytic_psd(f)
    src = _fisher_eq_source(name, 7)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for order, di in ((2, 0), (4, 1)):
        out = np.zeros((dim, dim))
        assert hh.hh_fisher_numerical(method.encode(), cases.DETECTORS[di].encode(), cases.DETECTORS[0].encode(), dim, order, C.byref(src),
                                      p(f), f.size, p(psd), p(out)) == 0
        _fisher_eq_check(out, oracle, method, src, f, np.tile(psd, (2, 1)), dim, order, di)


@pytest.mark.gpu
@pytest.mark.parametrize("case", FISHER_EQ_CASES, ids=[c[0] for c in FISHER_EQ_CASES])
def test_fisher_with_equatorial_orientation(ctx, oracle, case):
    name, method, kw, dim = case
    f = cases.grid(cases.FISHER_GRID)
    psd2 = np.tile(workloads.aligo_analytic_psd(f), (2, 1))
    ctx.set_network(cases.DETECTORS[:2], f, psd2)
    src = _fisher_eq_source(name, 7)
    for order, di in ((2, 0), (4, 1), (4, -1)):
        got = ctx.fisher_numerical_batch(method, [src], dim, order=order, detector_index=di)[0]
        _fisher_eq_check(got, oracle, method, src, f, psd2, dim, order, di)
    # a batch may mix both conventions (the flag is per source)
    plain = cases.source_from_bytes(np.load(os.path.join(GOLD, "fisher_v1.npz"))[name + "/src"])
    both = ctx.fisher_numerical_batch(method, [src, plain, src], dim, order=4, detector_index=1)
    alone = ctx.fisher_numerical_batch(method, [plain], dim, order=4, detector_index=1)[0]
    assert np.array_equal(both[1], alone) and np.array_equal(both[0], both[2])


@pytest.mark.gpu
def test_fisher_refuses_horizon_coordinates(ctx):
    f = cases.grid(cases.FISHER_GRID)
    ctx.set_network(cases.DETECTORS[:2], f, np.tile(workloads.aligo_analytic_psd(f), (2, 1)))
    s = _fisher_eq_source("F_D", 1)
    s.equatorial_orientation = 0
    s.horizon_coord = 1
    with pytest.raises(Exception, match="horizon_coord"):
        ctx.fisher_numerical_batch("IMRPhenomD", [s], 11)
