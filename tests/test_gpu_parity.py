"""GPU tier: the CUDA path, called through the C ABI, against (1) the committed golden vectors generated from the
reference's own code and (2) the oracle library itself when it travelled to the box, on identical inputs; plus
size-independent properties at BASELINE.json's full sizes.

Tolerances are BASELINE.json's: waveform <= 1e-10 relative to max|h|, log-likelihood <= 1e-9 relative.  Fisher matrices:
normalised measure e_ij = |dF_ij|/sqrt(F_ii F_jj): median(e) <= 1e-6 and max(e) <= max(1e-6, 3 x the reference's own
self-difference for that matrix: tests/fisher_noise.py, stored in tests/golden/fisher_noise_v2.npz).
"""
import os

import numpy as np
import pytest

import cases
from gw_analysis_tools_b200 import abi, workloads

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
WF_TOL = 1e-10
LL_TOL = 1e-9
FISHER_NORM_TOL = 1e-6
import fisher_noise  # noqa: E402

FISHER_NOISE_FACTOR = fisher_noise.FACTOR


@pytest.fixture(scope="module")
def gold_wf():
    return np.load(os.path.join(GOLD, "waveforms_v1.npz"))


@pytest.fixture(scope="module")
def gold_fisher():
    return np.load(os.path.join(GOLD, "fisher_v1.npz"))


@pytest.fixture(scope="module")
def gold_fisher_noise():
    return np.load(os.path.join(GOLD, "fisher_noise_v2.npz"))


@pytest.fixture(scope="module")
def gold_mcmc():
    return np.load(os.path.join(GOLD, "mcmc_v1.npz"))


def _relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _sources_from_oracle(oracle, wl, n):
    _, srcs = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params[:n], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd,
                                        None, return_sources=True)
    return srcs


def _inject(ctx, wl):
    """Zero-noise injection made by the product (as bench.py does)."""
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    wl.data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)
    return wl


# ---- every waveform family against the golden vectors ---------------------------------------------------------------------

@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_family_vs_golden(ctx, gold_wf, case):
    name, method, kw, gspec = case
    f = cases.grid(gspec)
    src = cases.source_from_bytes(gold_wf[name + "/src"])
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    resp_gold = gold_wf[name + "/resp"]
    data = cases.derived_data(resp_gold)
    ctx.set_network(cases.DETECTORS, f, psd, data)
    hp, hc = ctx.fourier_waveform_batch(method, [src])
    assert _relerr(hp[0], gold_wf[name + "/hp"]) <= WF_TOL
    assert _relerr(hc[0], gold_wf[name + "/hc"]) <= WF_TOL
    resp = ctx.coherent_response_batch(method, [src])[0]
    for d in range(3):
        assert _relerr(resp[d], resp_gold[d]) <= WF_TOL
    if name + "/single_L" in gold_wf:
        single = ctx.fourier_detector_response_batch(method, "Livingston", [src])[0]
        assert _relerr(single, gold_wf[name + "/single_L"]) <= WF_TOL
    ll = ctx.loglike_batch(method, [src])[0]
    ref = float(gold_wf[name + "/logL"])
    assert abs(ll - ref) <= LL_TOL * abs(ref), (ll, ref)


@pytest.mark.parametrize("cfg", [1, 2, 4, 5])
def test_mcmc_batch_vs_golden(ctx, gold_mcmc, cfg):
    """The bench's call path (sampling vectors -> logL) for small versions of the BASELINE configs."""
    L = 1024 if cfg != 5 else 4096
    wl = workloads.make(cfg, W=16, L=L)
    assert np.array_equal(wl.params, gold_mcmc["cfg%d/params" % cfg])
    ctx.set_network(wl.detectors, wl.f, wl.psd, gold_mcmc["cfg%d/data" % cfg])
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    ref = gold_mcmc["cfg%d/logL" % cfg]
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= LL_TOL, rel.max()


@pytest.mark.parametrize("case", cases.FISHER_CASES, ids=[c[0] for c in cases.FISHER_CASES])
def test_fisher_vs_golden(ctx, gold_fisher, gold_fisher_noise, case):
    name, method, kw, dim = case
    f = cases.grid(cases.FISHER_GRID)
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    src = cases.source_from_bytes(gold_fisher[name + "/src"])
    ctx.set_network(cases.DETECTORS, f, psd)
    for order in (2, 4):
        for d, det in enumerate(cases.DETECTORS[:2]):
            out = ctx.fisher_numerical_batch(method, [src], dim, order=order, detector_index=d, reference_index=0)[0]
            ref = gold_fisher["%s/o%d/%s" % (name, order, det)]
            dg = np.sqrt(np.abs(np.diag(ref)))
            nerr = np.abs(out - ref) / np.outer(dg, dg)
            assert np.median(nerr) <= FISHER_NORM_TOL, (name, order, det, np.median(nerr))
            floor = float(gold_fisher_noise["%s/o%d/%s" % (name, order, det)])
            assert nerr.max() <= max(FISHER_NORM_TOL, FISHER_NOISE_FACTOR * floor), (name, order, det, nerr.max(), floor)
            assert np.array_equal(out, out.T)
    total = ctx.fisher_numerical_batch(method, [src], dim, order=4, detector_index=-1, reference_index=0)[0]
    ref = gold_fisher[name + "/sum_o4"]
    dg = np.sqrt(np.abs(np.diag(ref)))
    assert (np.abs(total - ref) / np.outer(dg, dg)).max() <= max(FISHER_NORM_TOL, FISHER_NOISE_FACTOR * float(gold_fisher_noise[name + "/o4/sum"]))


def test_fisher_batch_is_consistent(ctx):
    """A batch of sources gives, source by source, exactly what single-source calls give (and survives chunking)."""
    srcs = workloads.fisher_sources(6)
    f = cases.grid((20.0, 0.5, 1024))
    ctx.set_network(cases.DETECTORS, f, np.tile(workloads.aligo_analytic_psd(f), (3, 1)))
    batch = ctx.fisher_numerical_batch("IMRPhenomD", srcs, 11, order=4)
    for i, s in enumerate(srcs):
        one = ctx.fisher_numerical_batch("IMRPhenomD", [s], 11, order=4)[0]
        assert np.array_equal(one, batch[i])
    assert np.all(np.isfinite(batch)) and np.all(np.diagonal(batch, axis1=1, axis2=2) > 0)


# ---- against the oracle library itself, on freshly drawn inputs --------------------------------------------------------------

@pytest.mark.parametrize("cfg,masses", [(1, (36.0, 29.0)), (1, (10.0, 8.0)), (2, (36.0, 29.0)), (2, (12.0, 7.0)), (4, (36.0, 29.0)),
                                        (5, (1.5, 1.3))])
def test_loglike_mcmc_vs_oracle(ctx, oracle, cfg, masses):
    L = 8192 if cfg != 5 else 1 << 15
    wl = _inject(ctx, workloads.make(cfg, W=48, L=L, masses=masses, seed=1234 + cfg))
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    ref = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data)
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= LL_TOL, rel.max()
    assert ctx.last_active_bins > 0


@pytest.mark.parametrize("cfg", [1, 2, 5])
def test_waveform_vs_oracle(ctx, oracle, cfg):
    L = 4096 if cfg != 5 else 1 << 14
    wl = workloads.make(cfg, W=6, L=L, seed=99 + cfg)
    srcs = _sources_from_oracle(oracle, wl, 6)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    hp, hc = ctx.fourier_waveform_batch(wl.method, srcs)
    resp = ctx.coherent_response_batch(wl.method, srcs)
    for w in range(6):
        rp, rc = oracle.fourier_waveform(wl.method, srcs[w], wl.f)
        assert _relerr(hp[w], rp) <= WF_TOL and _relerr(hc[w], rc) <= WF_TOL
        rr = oracle.coherent_response(wl.method, srcs[w], wl.detectors, wl.f)
        for d in range(wl.D):
            assert _relerr(resp[w, d], rr[d]) <= WF_TOL


def test_gaussleg_quadrature_vs_oracle(ctx, oracle):
    """GAUSSLEG integration on a log10-spaced, non-uniform grid (src/mcmc_gw.cpp:821-833)."""
    n = 600
    x, w = np.polynomial.legendre.leggauss(n)
    lo, hi = np.log10(20.0), np.log10(1000.0)
    logf = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
    wts = 0.5 * (hi - lo) * w
    f = 10 ** logf
    wl = workloads.make(1, W=24, L=64)
    wl.f, wl.psd = f, np.tile(workloads.aligo_analytic_psd(f), (2, 1))
    srcs = _sources_from_oracle(oracle, wl, 24)
    data = oracle.coherent_response(wl.method, srcs[0], wl.detectors, f)
    ctx.set_network(wl.detectors, f, wl.psd, data, weights=wts, integration_method="GAUSSLEG", log10F=True)
    got = ctx.loglike_batch(wl.method, srcs)
    ref = oracle.loglike_batch(wl.method, srcs, wl.detectors, f, wl.psd, data, weights=wts, integ="GAUSSLEG", log10F=True)
    assert (np.abs(got - ref) / np.abs(ref)).max() <= LL_TOL


def test_repack_matches_reference(ctx, oracle):
    for cfg in (1, 2, 4, 5):
        wl = workloads.make(cfg, W=16, L=1024)
        srcs = _sources_from_oracle(oracle, wl, 16)
        ctx.set_network(wl.detectors, wl.f, wl.psd)
        mine = ctx.repack_mcmc_batch(wl.method, wl.params[:16], wl.gmst, wl.mod)
        for a, b in zip(mine, srcs):
            for name in ("mass1", "mass2", "Luminosity_Distance", "RA", "DEC", "psi", "incl_angle", "phiRef"):
                assert abs(getattr(a, name) - getattr(b, name)) <= 4e-16 * max(1.0, abs(getattr(b, name))), name
            assert abs((wl.T_segment - a.tc) - b.tc) <= 1e-15 * wl.T_segment
            # the aligned-spin repack of the reference leaves the in-plane components uninitialised (src/fisher.cpp:2276-2277)
            for i in (range(3) if cfg == 2 else [2]):
                assert abs(a.spin1[i] - b.spin1[i]) <= 1e-15 and abs(a.spin2[i] - b.spin2[i]) <= 1e-15
            if cfg == 4:
                assert abs(a.betappe[0] - b.betappe[0]) <= 1e-15 * abs(b.betappe[0]) and a.Nmod == b.Nmod == 1
            if cfg == 5:
                assert abs(a.tidal_s - b.tidal_s) <= 1e-15 * b.tidal_s


def test_antenna_and_dtoa(ctx, oracle):
    rng = np.random.default_rng(5)
    W = 257
    RA, DEC, psi = rng.uniform(0, 2 * np.pi, W), np.arcsin(rng.uniform(-1, 1, W)), rng.uniform(0, np.pi, W)
    dets = ["Hanford", "Livingston", "Virgo", "Kagra", "Indigo", "CE", "ET1", "ET2"]
    f = 20 + np.arange(64.0)
    ctx.set_network(dets, f, np.ones((len(dets), 64)))
    fp, fc, dt = ctx.antenna_batch(RA, DEC, psi, 2.1)
    rfp, rfc, rdt = oracle.antenna_batch(RA, DEC, psi, 2.1, dets)
    assert np.abs(fp - rfp).max() <= 1e-14 and np.abs(fc - rfc).max() <= 1e-14
    assert np.abs(dt - rdt).max() <= 1e-16


# ---- size-independent properties at BASELINE's full sizes ----------------------------------------------------------------

def test_full_size_properties_cfg2(ctx):
    """IMRPhenomPv2, 3 detectors, 4096 walkers x 16384 bins (BASELINE configs[1])."""
    wl = _inject(ctx, workloads.make(2))
    W = wl.W
    full = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    assert full.shape == (W,) and np.all(np.isfinite(full))
    # (1) walkers are independent: any permutation / split of the batch gives bit-identical per-walker values
    perm = np.random.default_rng(0).permutation(W)
    assert np.array_equal(ctx.loglike_mcmc_batch(wl.method, wl.params[perm], wl.gmst, wl.T_segment, wl.mod), full[perm])
    half = ctx.loglike_mcmc_batch(wl.method, wl.params[:W // 2 + 7], wl.gmst, wl.T_segment, wl.mod)
    assert np.array_equal(half, full[:W // 2 + 7])
    # (2) the injected point maximises the zero-noise likelihood: logL(inj) = (d|d)/2 >= logL(anything else)
    ll_inj = ctx.loglike_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.T_segment, wl.mod)[0]
    assert ll_inj > 0 and np.all(full <= ll_inj * (1 + 1e-12))
    # (3) logL is quadratic in the amplitude scale a = DL_inj/DL:  logL(a) = a*DH - a^2*HH/2  -> third difference vanishes
    base = wl.params[:64].copy()
    a = np.array([0.5, 1.0, 1.5, 2.0])
    vals = []
    for s in a:
        p = base.copy()
        p[:, 6] = base[:, 6] - np.log(s)
        vals.append(ctx.loglike_mcmc_batch(wl.method, p, wl.gmst, wl.T_segment, wl.mod))
    v = np.array(vals)
    third = v[3] - 3 * v[2] + 3 * v[1] - v[0]
    scale = np.abs(v).max(axis=0)
    assert np.all(np.abs(third) <= 1e-9 * scale)
    # (4) an overall phase of the data rotates nothing in |.|^2 terms: HH part is invariant, checked through logL(d=0) = -HH/2 <= 0
    ctx.set_network(wl.detectors, wl.f, wl.psd, np.zeros_like(wl.data))
    hh_only = ctx.loglike_mcmc_batch(wl.method, wl.params[:256], wl.gmst, wl.T_segment, wl.mod)
    assert np.all(hh_only < 0)


def test_full_size_cfg5_long_grid(ctx):
    """IMRPhenomD_NRT on the 2^20-bin grid (BASELINE configs[4]); fewer walkers than the bench to keep the test short."""
    wl = _inject(ctx, workloads.make(5, W=64))
    assert wl.L == 1 << 20
    full = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    assert np.all(np.isfinite(full))
    ll_inj = ctx.loglike_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.T_segment, wl.mod)[0]
    assert ll_inj > 0 and np.all(full <= ll_inj * (1 + 1e-12))
    assert np.array_equal(ctx.loglike_mcmc_batch(wl.method, wl.params[::-1].copy(), wl.gmst, wl.T_segment, wl.mod), full[::-1])
    # the taper zeroes everything above 1.2 f_merger: those bins are skipped, so the active fraction is well below 1
    assert 0.1 < ctx.last_active_bins / (wl.W * wl.L) < 0.9


# ---- edge cases and error behaviour ---------------------------------------------------------------------------------------

def test_error_paths(ctx):
    from gw_analysis_tools_b200 import engine
    f = 20 + np.arange(64.0)
    with pytest.raises(engine.GwatB200Error):
        ctx.set_network(["Atlantis"], f, np.ones((1, 64)))
    ctx.set_network(["Hanford"], f, np.ones((1, 64)))
    with pytest.raises(engine.GwatB200Error):  # no data uploaded
        ctx.loglike_mcmc_batch("IMRPhenomD", np.zeros((1, 11)), 0.0, 1.0)
    ctx.set_network(["Hanford"], f, np.ones((1, 64)), np.zeros((1, 64), complex))
    with pytest.raises(engine.GwatB200Error):  # unknown method
        ctx.loglike_mcmc_batch("IMRPhenomXYZ", np.zeros((1, 11)), 0.0, 1.0)
    with pytest.raises(engine.GwatB200Error):  # wrong dimension
        ctx.loglike_mcmc_batch("IMRPhenomD", np.zeros((1, 12)), 0.0, 1.0)
    with pytest.raises(engine.GwatB200Error):  # GAUSSLEG without weights
        ctx.set_network(["Hanford"], f, np.ones((1, 64)), integration_method="GAUSSLEG")
    # an unphysical point (eta > 1/4) gives NaN, like the reference, and does not poison its neighbours
    wl = workloads.make(1, W=4, L=64)
    bad = wl.params.copy()
    bad[1, 8] = 0.3
    ctx.set_network(wl.detectors, wl.f, wl.psd, np.zeros((2, 64), complex))
    out = ctx.loglike_mcmc_batch(wl.method, bad, wl.gmst, wl.T_segment)
    assert np.isnan(out[1]) and np.all(np.isfinite(out[[0, 2, 3]]))
    assert ctx.loglike_mcmc_batch(wl.method, bad[:0], wl.gmst, wl.T_segment).size == 0


def test_odd_length_and_all_bins_above_cutoff(ctx, oracle):
    """Simpson's rule as the reference applies it for odd AND even L, and a grid that lies entirely above 0.2/M."""
    for L in (1001, 1002):
        wl = workloads.make(1, W=8, L=L)
        srcs = _sources_from_oracle(oracle, wl, 8)
        data = oracle.coherent_response(wl.method, srcs[3], wl.detectors, wl.f)
        ctx.set_network(wl.detectors, wl.f, wl.psd, data)
        got = ctx.loglike_batch(wl.method, srcs)
        ref = oracle.loglike_batch(wl.method, srcs, wl.detectors, wl.f, wl.psd, data)
        assert (np.abs(got - ref) / np.abs(ref)).max() <= LL_TOL
    heavy = abi.source_defaults(**dict(cases.BBH, mass1=400.0, mass2=300.0))
    f = 100.0 + np.arange(512.0)
    ctx.set_network(cases.DETECTORS, f, np.ones((3, 512)), np.ones((3, 512), complex))
    assert ctx.loglike_batch("IMRPhenomD", [heavy])[0] == 0.0
    hp, hc = ctx.fourier_waveform_batch("IMRPhenomD", [heavy])
    assert not hp.any() and not hc.any()
    assert ctx.last_active_bins == 0


def test_glq_grid_end_to_end(ctx, oracle):
    """The library's own Gauss-Legendre grid (log10 f nodes) through the GAUSSLEG likelihood, against the reference fed the same grid."""
    from gw_analysis_tools_b200 import engine
    f, w = engine.gauss_legendre_grid(20.0, 1024.0, 300, True)
    wl = workloads.make(2, W=24, L=1024)
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    ctx.set_network(wl.detectors, f, psd, None, w, "GAUSSLEG", True)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, f, psd, data, w, "GAUSSLEG", True)
    got = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    want = oracle.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, f, psd, data, weights=w,
                                     integ="GAUSSLEG", log10F=True)
    assert (np.abs(got - want) / np.abs(want)).max() <= LL_TOL


def test_concurrent_callers_share_a_context(ctx, gold_mcmc):
    """Thread safety of the C ABI ('Threading'): several host threads submit batches to ONE context at once
    (ctypes releases the GIL); every call returns exactly what it returns alone."""
    import threading
    wl = workloads.make(2, W=64, L=1024)
    ctx.set_network(wl.detectors, wl.f, wl.psd, gold_mcmc["cfg2/data"])
    chunks = [wl.params[i::4] for i in range(4)]
    alone = [ctx.loglike_mcmc_batch(wl.method, c, wl.gmst, wl.T_segment, wl.mod) for c in chunks]
    results = [[None] * 8 for _ in range(4)]

    def work(k):
        for it in range(8):
            results[k][it] = ctx.loglike_mcmc_batch(wl.method, chunks[k], wl.gmst, wl.T_segment, wl.mod)
    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(4):
        for it in range(8):
            assert np.array_equal(results[k][it], alone[k])


@pytest.mark.parametrize("cfg,L", [(1, 8192), (2, 16384), (4, 3000), (5, 70000)])
def test_value_does_not_depend_on_the_batch(ctx, cfg, L):
    """The fixed summation tree: a walker's logL is bit-identical alone, in a small group and in a large batch,
    whatever run length per CTA the launcher picks for the ensemble size."""
    wl = workloads.make(cfg, W=600, L=L)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    ctx.set_network(wl.detectors, wl.f, wl.psd, ctx.coherent_response_batch(wl.method, src)[0])
    full = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    assert np.all(np.isfinite(full))
    for i in (0, 17, 599):
        assert ctx.loglike_mcmc_batch(wl.method, wl.params[i:i + 1], wl.gmst, wl.T_segment, wl.mod)[0] == full[i]
    part = np.concatenate([ctx.loglike_mcmc_batch(wl.method, wl.params[a:a + 37], wl.gmst, wl.T_segment, wl.mod) for a in range(0, 600, 37)])
    assert np.array_equal(part, full)
