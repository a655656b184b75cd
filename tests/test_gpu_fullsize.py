"""GPU tier, BASELINE.json's REAL shapes against the oracle (the reference's own code, oracle/_ref) on identical inputs.

tests/test_gpu_parity.py compares with the oracle on shortened grids; here nothing is shortened along the bin axis:
  cfg1  IMRPhenomD      2 det   8192 bins   (36,29) and (10,8) Msun
  cfg2  IMRPhenomPv2    3 det  16384 bins   (36,29) and (10,8) Msun  -- the light set keeps every bin below 0.2/M active
  cfg4  dCS_IMRPhenomD  3 det  16384 bins   both mass sets
  cfg5  IMRPhenomD_NRT  3 det   2^20 bins   10 Hz .. 4106 Hz at df = 1/256 Hz: inspiral, merger and the Planck-taper band
  cfg5 with t_c in the middle of the 256 s segment (carrier argument 2 pi t_c f ~ 3e6 rad)
  cfg3  the Fisher bench population (m in U(3,100) Msun, workloads.fisher_sources), 64 sources, 4096 bins, order 4
The walker counts are what the CPU oracle finishes in seconds; per-walker values do not depend on the batch
(test_value_does_not_depend_on_the_batch), so these are the very numbers the full ensembles produce.
Tolerances: BASELINE.json's (waveform 1e-10 of max|h|, logL 1e-9 relative; Fisher: see test_fisher_bench_population).
"""
import os

import numpy as np
import pytest

from gw_analysis_tools_b200 import workloads

pytestmark = pytest.mark.gpu

WF_TOL = 1e-10
LL_TOL = 1e-9


def _inject(ctx, wl):
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    wl.data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)
    return wl


def _threads():
    return max(1, len(os.sched_getaffinity(0)))


@pytest.mark.parametrize("cfg,masses,W", [(1, (36.0, 29.0), 64), (1, (10.0, 8.0), 64), (2, (36.0, 29.0), 64), (2, (10.0, 8.0), 64),
                                          (4, (36.0, 29.0), 64), (4, (10.0, 8.0), 64)])
def test_loglike_full_grid_vs_oracle(ctx, oracle, cfg, masses, W):
    wl = _inject(ctx, workloads.make(cfg, W=W, masses=masses, seed=4321 + cfg))
    assert wl.L == (8192 if cfg == 1 else 16384)
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    frac = ctx.last_active_bins / (W * wl.L)
    ref = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data,
                                    nthreads=_threads())
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= LL_TOL, (cfg, masses, rel.max())
    # the light set reaches the top of the grid (0.2/M = 2256 Hz for 18 Msun): (almost) every bin is evaluated and compared
    if masses[0] < 20:
        assert frac > 0.95, frac
    else:
        assert 0.1 < frac < 0.9, frac


def test_cfg5_full_grid_vs_oracle(ctx, oracle):
    """IMRPhenomD_NRT on the whole 2^20-bin grid: logL of 8 walkers and the waveform of 2 of them bin by bin."""
    wl = _inject(ctx, workloads.make(5, W=8, seed=77))
    assert wl.L == 1 << 20
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    ref, srcs = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data,
                                          nthreads=_threads(), return_sources=True)
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= LL_TOL, rel.max()
    # waveform and responses over all bins: covers the merger and the taper band (f_merger ~ 1.5-2 kHz; zero above 1.2 f_merger)
    two = [srcs[0], srcs[5]]
    hp, hc = ctx.fourier_waveform_batch(wl.method, two)
    resp = ctx.coherent_response_batch(wl.method, two)
    for w in range(2):
        rp, rc = oracle.fourier_waveform(wl.method, two[w], wl.f)
        live = np.flatnonzero(np.abs(rp) > 0)
        assert live.size and wl.f[live[-1]] > 1000.0           # the model reaches the kHz band ...
        assert wl.f[live[-1]] < wl.f[-1]                       # ... and the taper cuts it off inside the grid
        assert np.abs(hp[w] - rp).max() <= WF_TOL * np.abs(rp).max()
        assert np.abs(hc[w] - rc).max() <= WF_TOL * np.abs(rc).max()
        # the taper band itself (the last 20 % of the live range), relative to its own maximum
        band = slice(live[int(0.8 * live.size)], live[-1] + 1)
        assert np.abs(hp[w][band] - rp[band]).max() <= 1e-9 * np.abs(rp[band]).max()
        assert not hp[w][live[-1] + 1:].any()
        rr = oracle.coherent_response(wl.method, two[w], wl.detectors, wl.f)
        for d in range(wl.D):
            assert np.abs(resp[w, d] - rr[d]).max() <= WF_TOL * np.abs(rr[d]).max()


def test_cfg5_mid_segment_coalescence_vs_oracle(ctx, oracle):
    """t_c = T/2 = 128 s on the 2^20-bin grid: the time-shift term of the carrier phase reaches 2 pi 128 s x 4 kHz = 3e6 rad."""
    wl = workloads.make(5, W=8, seed=78)
    wl.inj[5] = 128.0
    wl.params[:, 5] = 128.0 + 1e-3 * np.random.default_rng(3).standard_normal(wl.W)
    _inject(ctx, wl)
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    ref = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data,
                                    nthreads=_threads())
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= LL_TOL, rel.max()


def test_cfg2_mid_segment_coalescence_vs_oracle(ctx, oracle):
    wl = workloads.make(2, W=32, masses=(10.0, 8.0), seed=79)
    wl.inj[5] = 4.0
    wl.params[:, 5] = 4.0 + 1e-3 * np.random.default_rng(4).standard_normal(wl.W)
    _inject(ctx, wl)
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    ref = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data,
                                    nthreads=_threads())
    assert (np.abs(got - ref) / np.abs(ref)).max() <= LL_TOL


def test_fisher_bench_population(ctx, oracle):
    """cfg3's own population (bench.py --config 3): 64 sources with m in U(3,100) Msun, 3 detectors summed, 4096 bins, order 4.

    Measure: e_ij = |dF_ij| / sqrt(F_ii F_jj).  The eps = 1e-8 stencil amplifies rounding by 1e8, so the reference differs from
    ITSELF when its arithmetic is perturbed; the yardstick is measured here, per source, by running the reference again on
    inputs moved by parts in 1e14 (far below the stencil step, far above nothing: a different rounding sequence, the same
    mathematics).  The GPU must agree with the reference to within FACTOR x that self-difference, for the largest entry and
    for the median entry of every matrix (or to BASELINE.json's 1e-6 where the reference is quieter than that)."""
    import fisher_noise
    S = 64
    srcs = workloads.fisher_sources(S)
    f = 20.0 + 0.25 * np.arange(4096)
    dets = ["Hanford", "Livingston", "Virgo"]
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    ctx.set_network(dets, f, psd)
    got = ctx.fisher_numerical_batch("IMRPhenomD", srcs, 11, order=4)
    ref = oracle.fisher_numerical_batch("IMRPhenomD", srcs, dets, f, psd, 11, order=4, detector_index=-1, reference_index=0,
                                        nthreads=_threads())
    floor, floor_med = fisher_noise.reference_self_difference(oracle, "IMRPhenomD", srcs, dets, f, psd, 11, 4, nthreads=_threads(),
                                                              with_median=True)
    finite = np.all(np.isfinite(ref.reshape(S, -1)), axis=1)
    assert finite.sum() >= S - 2
    assert np.array_equal(np.all(np.isfinite(got.reshape(S, -1)), axis=1), finite)  # NaN where, and only where, the reference has NaN
    worst = 0.0
    for i in np.flatnonzero(finite):
        e = fisher_noise.normalised_error(got[i], ref[i])
        assert np.median(e) <= max(1e-6, fisher_noise.FACTOR * floor_med[i]), (i, np.median(e), floor_med[i])
        bound = max(1e-6, fisher_noise.FACTOR * floor[i])
        worst = max(worst, e.max() / bound)
        assert e.max() <= bound, (i, e.max(), floor[i])
    print("fisher bench population: worst max-error / bound = %.2f" % worst)
