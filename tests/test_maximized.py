"""tc/phic-maximised likelihoods: gwat_b200_loglike_maximized_batch against golden values produced by the
reference's own maximized_Log_Likelihood_{aligned,unaligned}_spin_internal (oracle/_ref; the FFT there is the textbook
transform of standins/fftw3.h, cuFFT here), and against the oracle itself on fresh draws.  Tolerance: 1e-9 relative,
the log-likelihood tolerance of BASELINE.json.
"""
import os

import numpy as np
import pytest

import cases
from gw_analysis_tools_b200 import abi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 1e-9


def _case_inputs(name, gspec):
    gold = np.load(os.path.join(GOLD, "waveforms_v1.npz"))
    f = cases.grid(gspec)
    return f, np.tile(workloads.aligo_analytic_psd(f), (3, 1)), cases.maximized_data(gold, gspec), cases.source_from_bytes(gold[name + "/src"])


def test_oracle_reproduces_golden_maximized(oracle):
    want = np.load(os.path.join(GOLD, "maximized_v1.npz"))
    for name, method, kw, gspec in cases.CASES[::3]:
        f, psd, data, src = _case_inputs(name, gspec)
        got = oracle.loglike_maximized_batch(method, [src], cases.DETECTORS, f, psd, data)[0]
        assert got == float(want[name]), (name, got, float(want[name]))


def test_oracle_fft_stand_in_matches_numpy(oracle):
    """The reference's aligned-spin formula recomputed with numpy's FFT from the reference's own waveform."""
    name, method, kw, gspec = cases.CASES[0]
    f, psd, data, src = _case_inputs(name, gspec)
    got = oracle.loglike_maximized_batch(method, [src], cases.DETECTORS, f, psd, data)[0]
    frame = cases.source_from_bytes(np.load(os.path.join(GOLD, "waveforms_v1.npz"))[name + "/src"])
    frame.psi, frame.phiRef, frame.f_ref, frame.incl_angle, frame.tc = 0, 1, 10, 0, 1
    hp, _ = oracle.fourier_waveform(method, frame, f)
    df = f[1] - f[0]
    coef = np.where(np.arange(f.size) % 2 == 0, 2.0, 4.0)
    coef[0] = coef[-1] = 1
    total = 0
    for d in range(3):
        HH = 4 * (coef * np.abs(hp) ** 2 / psd[d]).sum() * df / 3
        G = np.fft.fft(4 * np.conj(data[d]) * hp / psd[d])
        total += 0.5 * (np.abs(G) ** 2).max() * df * df / HH
    assert abs(got - total) <= 1e-12 * abs(total)


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_maximized_vs_golden(ctx, case):
    name, method, kw, gspec = case
    want = float(np.load(os.path.join(GOLD, "maximized_v1.npz"))[name])
    f, psd, data, src = _case_inputs(name, gspec)
    ctx.set_network(cases.DETECTORS, f, psd, data)
    got = ctx.loglike_maximized_batch(method, [src])[0]
    assert abs(got - want) <= TOL * abs(want), (name, got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,L", [(1, 2048), (2, 2048), (2, 1000), (5, 4096)])
def test_maximized_batch_vs_oracle(ctx, oracle, cfg, L):
    """Fresh draws, whole batches (chunking included), power-of-two and other lengths."""
    wl = workloads.make(cfg, W=48, L=L)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, np.vstack([wl.inj[None, :], wl.params]), wl.gmst, wl.mod)
    inj = src[0]
    inj.tc = wl.T_segment - inj.tc
    data = ctx.coherent_response_batch(wl.method, [inj])[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    walkers = [src[i] for i in range(1, 49)]
    got = ctx.loglike_maximized_batch(wl.method, walkers)
    want = oracle.loglike_maximized_batch(wl.method, walkers, wl.detectors, wl.f, wl.psd, data)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() <= TOL, rel.max()


@pytest.mark.gpu
def test_maximized_properties(ctx):
    """The maximisation removes tc, phiRef, psi, inclination and the distance scale; the injected parameters score above the
    typical neighbour (not above all of them: the time axis is sampled at 1/(L df), coarser than the signal's bandwidth, as in
    the reference); a non-uniform grid is refused (the time axis is an FFT)."""
    wl = workloads.make(1, W=16, L=4096)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, np.vstack([wl.inj[None, :], wl.params]), wl.gmst, wl.mod)
    inj = src[0]
    inj.tc = wl.T_segment - inj.tc
    data = ctx.coherent_response_batch(wl.method, [inj])[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    base = ctx.loglike_maximized_batch(wl.method, [src[i] for i in range(17)])
    moved = []
    for i in range(17):
        s = abi.Source.from_buffer_copy(src[i])
        s.tc, s.phiRef, s.psi, s.incl_angle = 3.3, 0.4, 1.0, 1.2
        s.Luminosity_Distance = 3 * s.Luminosity_Distance
        moved.append(s)
    again = ctx.loglike_maximized_batch(wl.method, moved)
    assert np.allclose(again, base, rtol=1e-12)
    assert base[0] >= np.median(base[1:])
    f2 = wl.f.copy()
    f2[10:] += 0.01
    ctx.set_network(wl.detectors, f2, wl.psd, data)
    with pytest.raises(Exception):
        ctx.loglike_maximized_batch(wl.method, [src[0]])
