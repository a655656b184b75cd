"""Noise curves and SNR (input preparation; gwatpy's populate_noise_py / calculate_snr_py).

CPU tier: gwat_b200_populate_noise against the reference's populate_noise (oracle) -- analytic models everywhere, tabulated
curves read from the reference's own CSV files when the reference tree is on this box -- and against committed golden values.
GPU tier: gwat_b200_snr_batch against the reference's calculate_snr for several families, and against SNR^2 = -2 logL(data = 0).
"""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from gw_analysis_tools_b200 import abi, engine, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOISE_DIR = "/root/reference/data/noise_data/currently_supported"
CURVES = ["AdLIGODesign", "AdLIGOAPlus_smoothed", "CE1", "CE2_smoothed", "AdVIRGOPlus2_opt", "AdVIRGOPlus1_smoothed", "KAGRA_opt",
          "KAGRA_pess", "ET-D", "ET-D_smoothed", "AdLIGOVoyager", "AdLIGOMidHigh"]
# sqrt(S_n) at 25, 100 and 1000 Hz as the reference's populate_noise returns them (oracle build, 17 significant digits)
GOLD_ANALYTIC = {
    "aLIGO_analytic": [6.89110731305209e-24, 1.3899856114363197e-24, 1.0981322352066713e-24],
    "Hanford_O1_fitted": [8.204704326395266e-23, 8.7208359049412e-24, 1.8648899991438415e-23],
}


def test_analytic_curves_golden_and_oracle(oracle):
    f = np.array([25.0, 100.0, 1000.0])
    for name, want in GOLD_ANALYTIC.items():
        got = engine.populate_noise(f, name)
        assert np.allclose(got, want, rtol=1e-14, atol=0), (name, got)
    f = np.geomspace(5.0, 4000.0, 513)
    for name in GOLD_ANALYTIC:
        assert np.allclose(engine.populate_noise(f, name), oracle.populate_noise(f, name), rtol=1e-14, atol=0)
    assert np.allclose(engine.populate_noise(f, "aLIGO_analytic") ** 2, workloads.aligo_analytic_psd(f), rtol=1e-14, atol=0)


def test_gps_to_gmst_vs_reference(oracle):
    for gps in (1126259462.4, 1187008882.4, 630763213.0, 1e9 + 0.5, 1.4e9, 815000000.25):
        assert engine.gps_to_gmst_radian(gps) == oracle.gps_to_gmst_radian(gps)
    assert engine.gps_to_gmst_radian(1126259462.4) == 2.4568247373045096  # value of the reference build, GW150914's GPS time


def test_unknown_and_unsupported_curves():
    f = np.array([30.0, 40.0])
    for name, code in (("LISA", abi.ERR_UNSUPPORTED), ("LISA_SADC_CONF", abi.ERR_UNSUPPORTED), ("NoSuchCurve", abi.ERR_ARG),
                       ("CE1", abi.ERR_ARG)):  # tabulated curve without a directory
        with pytest.raises(engine.GwatB200Error) as e:
            engine.populate_noise(f, name)
        assert e.value.code == code
    with pytest.raises(engine.GwatB200Error) as e:
        engine.populate_noise(f, "CE1", "/nonexistent")
    assert e.value.code == abi.ERR_STATE


@pytest.mark.skipif(not os.path.isdir(NOISE_DIR), reason="the reference's noise tables are not on this box")
@pytest.mark.parametrize("curve", CURVES)
def test_tabulated_curves_vs_reference(oracle, curve):
    tab = np.loadtxt(os.path.join(NOISE_DIR, {"AdLIGODesign": "aligo_O4high.csv", "KAGRA_opt": "kagra_128Mpc.csv"}.get(curve, "")), delimiter=",") \
        if curve in ("AdLIGODesign", "KAGRA_opt") else None
    rng = np.random.default_rng(11)
    lo, hi = (12.0, 1800.0) if "KAGRA" not in curve else (12.0, 1500.0)
    f = np.sort(rng.uniform(lo, hi, 2000))
    got = engine.populate_noise(f, curve, NOISE_DIR)
    want = oracle.populate_noise(f, curve)
    assert np.all(np.isfinite(want)) and np.all(want > 0)
    assert np.array_equal(got, want)
    if tab is not None:  # exact at the knots
        k = tab[(tab[:, 0] > lo) & (tab[:, 0] < hi)][::37]
        assert np.array_equal(engine.populate_noise(k[:, 0], curve, NOISE_DIR), k[:, 1])


@pytest.mark.skipif(not os.path.isdir(NOISE_DIR), reason="the reference's noise tables are not on this box")
def test_outside_the_table_is_nan_and_an_error():
    lib = engine.load_library()
    f = np.array([1e-3, 100.0])
    out = np.zeros(2)
    dp = C.POINTER(C.c_double)
    rc = lib.gwat_b200_populate_noise(f.ctypes.data_as(dp), b"AdLIGODesign", NOISE_DIR.encode(), 2, out.ctypes.data_as(dp))
    assert rc == abi.ERR_ARG and np.isnan(out[0]) and out[1] > 0


SNR_CASES = [c for c in cases.CASES if c[0] in ("D_bbh", "P_full", "dCS", "NRT_love", "gIMR_log", "ppE_imr")]


@pytest.mark.gpu
@pytest.mark.parametrize("case", SNR_CASES, ids=[c[0] for c in SNR_CASES])
def test_snr_vs_reference(ctx, oracle, case):
    name, method, kw, gspec = case
    f = cases.grid(gspec)
    src = cases.source(kw)
    for det, curve in (("Hanford", "aLIGO_analytic"), ("Virgo", "Hanford_O1_fitted")):
        psd = engine.populate_noise(f, curve) ** 2
        ctx.set_network([det], f, psd[None, :])
        got = ctx.snr_batch(method, [src, src])
        want = oracle.calculate_snr(curve, det, method, src, f)
        assert got[0] == got[1]
        assert abs(got[0] - want) <= 1e-9 * want, (got[0], want)


@pytest.mark.gpu
def test_snr_ignores_data_and_matches_the_likelihood_of_empty_data(ctx):
    wl = workloads.make(2, W=32, L=2048)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.params, wl.gmst, wl.mod)
    snr = ctx.snr_batch(wl.method, src)
    data = ctx.coherent_response_batch(wl.method, src[:1])[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    assert np.array_equal(ctx.snr_batch(wl.method, src), snr)  # data present: ignored
    ctx.set_network(wl.detectors, wl.f, wl.psd, np.zeros_like(data))
    ll = ctx.loglike_batch(wl.method, src)
    assert np.allclose(snr ** 2, -2.0 * ll, rtol=1e-13, atol=0)

@pytest.mark.gpu
def test_snr_gauss_legendre_vs_reference(ctx, oracle):
    """calculate_snr with integration_method "GAUSSLEG" on log10-frequency nodes (src/waveform_util.cpp:479-510)."""
    fg, wg = engine.gauss_legendre_grid(20.0, 1024.0, 512, True)
    psd = engine.populate_noise(fg, "aLIGO_analytic") ** 2
    for name, method, kw, gspec in SNR_CASES[:3]:
        src = cases.source(kw)
        ctx.set_network(["Livingston"], fg, psd[None, :], None, wg, "GAUSSLEG", True)
        got = ctx.snr_batch(method, [src])[0]
        want = oracle.calculate_snr("aLIGO_analytic", "Livingston", method, src, fg, weights=wg, integ="GAUSSLEG", log10F=True)
        assert abs(got - want) <= 1e-9 * want, (name, got, want)
