"""LOSC text files -> frequency-domain data: gwat_b200_losc_prepare against the reference's allocate_LOSC_data.

The reference ships no LOSC strain files, so the inputs are synthetic files in the LOSC layout (three header lines, one sample
per line; PSD file with a header line and rows "f S_1 ... S_D").  CPU tier: argument/IO errors need no GPU work... but a context
does, so everything here is GPU tier except the oracle's own self-check.
"""
import os

import numpy as np
import pytest

from gw_analysis_tools_b200 import abi, engine


def write_losc(tmp_path, D=2, fs=256, duration=16, start=1126259446, df=0.25, fmin=20.0, fmax=100.0, seed=5):
    rng = np.random.default_rng(seed)
    n = fs * duration
    t = np.arange(n) / fs
    files = []
    for d in range(D):
        x = 1e-21 * (rng.standard_normal(n) + 3.0 * np.sin(2 * np.pi * (30.0 + 7 * d) * t + 0.3 * d) * np.exp(-((t - 9.0) / 2.0) ** 2))
        p = tmp_path / ("strain_%d.txt" % d)
        with open(p, "w") as fh:
            fh.write("# Gravitational wave strain for detector %d\n# This file has %d samples per second\n# starting GPS %d duration %d\n"
                     % (d, fs, start, duration))
            fh.write("\n".join("%.17e" % v for v in x) + "\n")
        files.append(str(p))
    f = np.arange(fmin, fmax + 0.5 * df, df)
    psd = np.stack([1e-46 * (1 + (50.0 / f) ** 4) * (1 + 0.1 * d) for d in range(D)], axis=1)
    pfile = tmp_path / "psd.txt"
    with open(pfile, "w") as fh:
        fh.write("# f " + " ".join("psd%d" % d for d in range(D)) + "\n")
        for i in range(f.size):
            fh.write("%.17e " % f[i] + " ".join("%.17e" % v for v in psd[i]) + "\n")
    return files, str(pfile), f, psd.T, n


@pytest.mark.gpu
@pytest.mark.parametrize("D,post", [(2, 1.0), (3, 2.0), (1, 0.5)])
def test_losc_prepare_vs_reference(ctx, oracle, tmp_path, D, post):
    files, pfile, f, psd, nsamp = write_losc(tmp_path, D=D)
    trigger = 1126259446 + 9.0
    gf, gpsd, gdata = ctx.losc_prepare(files, pfile, trigger, post)
    rf, rpsd, rdata = oracle.losc(files, pfile, trigger, post, f.size, nsamp)
    assert np.array_equal(gf, f) and np.array_equal(gf, rf)
    assert np.array_equal(gpsd, psd) and np.array_equal(gpsd, rpsd)
    scale = np.abs(rdata).max()
    assert scale > 0 and np.all(np.isfinite(gdata))
    assert np.abs(gdata - rdata).max() <= 1e-10 * scale
    # and against numpy directly: the Tukey-windowed segment's DFT times dt
    fs, Tobs = 256, 4.0
    n = int(Tobs * fs)
    x = np.loadtxt(files[0], skiprows=3)
    tt = 1126259446 + np.arange(x.size) / fs
    sel = (tt > trigger - (Tobs - post)) & (tt <= trigger + post)
    alpha = 0.8 / Tobs
    imin, imax = int(alpha * (n - 1) / 2.0), int((n - 1) * (1 - alpha / 2.0))
    i = np.arange(n)
    w = np.where(i < imin, 0.5 * (1 + np.cos(np.pi * (i / imin - 1))), np.where(i < imax, 1.0, 0.5 * (1 + np.cos(np.pi * (i / imin - 2 / alpha + 1)))))
    spec = np.fft.fft(x[sel][:n] * w) / fs
    k0 = int(round(f[0] * Tobs))
    assert np.abs(gdata[0] - spec[k0:k0 + f.size]).max() <= 1e-10 * scale


@pytest.mark.gpu
def test_losc_prepare_feeds_the_likelihood(ctx, tmp_path):
    """The prepared arrays go straight into set_network; the likelihood of a template against them is finite."""
    from gw_analysis_tools_b200 import workloads
    files, pfile, f, psd, _ = write_losc(tmp_path, D=2)
    gf, gpsd, gdata = ctx.losc_prepare(files, pfile, 1126259446 + 9.0, 1.0)
    ctx.set_network(["Hanford", "Livingston"], gf, gpsd, gdata)
    wl = workloads.make(1, W=8, L=1024)
    ll = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, 4.0, wl.mod)
    assert np.all(np.isfinite(ll))


@pytest.mark.gpu
def test_losc_prepare_errors(ctx, tmp_path):
    files, pfile, f, psd, _ = write_losc(tmp_path, D=2)
    with pytest.raises(engine.GwatB200Error) as e:
        ctx.losc_prepare(files, str(tmp_path / "missing.txt"), 1126259446 + 9.0, 1.0)
    assert e.value.code == abi.ERR_STATE
    with pytest.raises(engine.GwatB200Error) as e:
        ctx.losc_prepare([files[0], str(tmp_path / "nope.txt")], pfile, 1126259446 + 9.0, 1.0)
    assert e.value.code == abi.ERR_STATE
    with pytest.raises(engine.GwatB200Error) as e:  # the reference prints an error and leaves its outputs unallocated here
        ctx.losc_prepare(files, pfile, 1126259446 + 15.5, 1.0)
    assert e.value.code == abi.ERR_ARG and "trigger" in str(e.value)
