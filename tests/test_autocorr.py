"""Autocorrelation lengths for the thinned chain output: calc_ac_vals -> auto_corr_from_data_batch ->
auto_correlation_spectral_windowed (src/mcmc_io_util.cpp:434-520, src/autocorrelation.cpp:152-347, 401-462).

CPU tier: the compiled reference against a numpy statement of the estimator (pins what the device code has to reproduce).
GPU tier: gwat_b200_autocorrelation_lengths (batched cuFFT) against the compiled reference."""
import numpy as np
import pytest

from gw_analysis_tools_b200 import sampler


def ar1(n, phi, rng):
    x = np.zeros(n)
    e = rng.standard_normal(n)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    return x


def chains(steps, seed, n_chains=3):
    """AR(1) chains with integrated autocorrelation times from ~3 to ~65 steps, on different offsets and scales, and one trend."""
    rng = np.random.default_rng(seed)
    phis = (0.5, 0.9, 0.97, 0.0)
    pos = np.zeros((n_chains, steps, len(phis) + 1))
    for c in range(n_chains):
        for d, phi in enumerate(phis):
            pos[c, :, d] = (c + 1) * 10.0 ** (d - 1) * (3.0 + ar1(steps, phi, rng))
        pos[c, :, -1] = np.arange(steps) * 1e-3 + 0.01 * rng.standard_normal(steps)  # a drifting parameter: tau of the order of the length
    return pos


def windowed_tau(x):
    n = x.size
    L = 2 * 2 ** int(np.ceil(np.log2(n)))
    y = np.zeros(L)
    y[:n] = x - x.mean()
    acov = np.fft.fft(np.abs(np.fft.fft(y)) ** 2).real  # (the unnormalised backward transform of a real, even sequence)
    t = 2 * np.cumsum(acov[:n] / acov[0]) - 1
    hit = np.nonzero(np.arange(n) > 5 * t)[0]
    return t[hit[0] if hit.size else n - 1]


@pytest.mark.parametrize("steps,begin", [(700, 0), (1100, 76)])
def test_reference_estimator_is_the_windowed_emcee_one(oracle, steps, begin):
    pos = chains(steps, 3)
    ac, tau = oracle.autocorrelation_lengths(pos, begin=begin)
    mine = np.array([[windowed_tau(pos[c, begin:, d]) for d in range(pos.shape[2])] for c in range(pos.shape[0])])
    assert np.abs(tau - mine).max() <= 1e-9 * np.abs(mine).max()
    assert np.array_equal(ac, tau.astype(np.int32))
    assert 2 <= ac[:, 0].min() and ac[:, 0].max() <= 5 and ac[:, 2].min() >= 10  # phi = 0.5 -> ~3, phi = 0.97 -> tens of steps


@pytest.mark.gpu
@pytest.mark.parametrize("steps,begin", [(700, 0), (1024, 0), (1025, 1), (3000, 123), (4099, 3)])
def test_autocorrelation_lengths_vs_reference(ctx, oracle, steps, begin):
    pos = chains(steps, 11 + steps)
    ref_ac, ref_tau = oracle.autocorrelation_lengths(pos, begin=begin)
    ac, tau = sampler.autocorrelation_lengths(ctx, pos, begin=begin)
    assert np.abs(tau - ref_tau).max() <= 1e-9 * np.abs(ref_tau).max(), (tau, ref_tau)
    # the lag is the truncated estimator: equal unless the estimator sits within rounding of an integer
    near_integer = np.abs(ref_tau - np.round(ref_tau)) < 1e-9 * np.abs(ref_tau)
    assert np.array_equal(ac[~near_integer], ref_ac[~near_integer])
    # the same rows in another batch, and alone
    one_ac, one_tau = sampler.autocorrelation_lengths(ctx, pos[1:2], begin=begin)
    assert np.array_equal(one_ac[0], ac[1]) and np.array_equal(one_tau[0], tau[1])


@pytest.mark.gpu
def test_autocorrelation_short_rows_and_errors(ctx, oracle):
    pos = chains(2, 5)
    ac, tau = sampler.autocorrelation_lengths(ctx, pos)
    ref_ac, _ = oracle.autocorrelation_lengths(pos)
    assert np.array_equal(ac, ref_ac) and np.all(ac == 2)  # the brute-force branch of rows <= MAX_SERIAL leaves at lag 2
    from gw_analysis_tools_b200 import engine
    with pytest.raises(engine.GwatB200Error):
        sampler.autocorrelation_lengths(ctx, chains(10, 1), begin=10)


@pytest.mark.gpu
def test_thinned_output_with_device_autocorrelation(ctx, tmp_path):
    """The lags feed write_flat_thin_output exactly as calc_ac_vals' feed the reference's (src/mcmc_io_util.cpp:555-642)."""
    from gw_analysis_tools_b200 import chain_io
    pos = chains(2000, 9)[:, :, :4]
    ac, _ = sampler.autocorrelation_lengths(ctx, pos, begin=100)
    path = str(tmp_path / "thin.gwat")
    trim = np.full(pos.shape[0], 100, dtype=np.int32)
    rows = chain_io.write_flat_thin_output(path, pos, ac, trim)
    # count_indep_samples (src/mcmc_io_util.cpp:521-552): mean length over the mean of the chains' largest lags
    assert rows == int((2000 - 100) / np.mean(ac.max(axis=1))) and rows > 10
