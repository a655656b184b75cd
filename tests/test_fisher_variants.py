"""Pointed (non-sky-averaged) Fisher matrices for the families the golden Fisher cases do not hold: NRTidal (one and two tidal
parameters, physical and MCMC_ sets), gIMR, dCS through its ppE mapping, ppE IMR, precessing ppE -- unpack_parameters / repack_parameters
(src/fisher.cpp:1841-2507) and the response stencil (:340-557) against the compiled reference at its own noise floor."""
import ctypes as C

import numpy as np
import pytest

import fisher_noise
from gw_analysis_tools_b200 import abi, workloads


def _with(src, **kw):
    t = abi.Source()
    C.memmove(C.addressof(t), C.addressof(src), C.sizeof(src))
    for k, v in kw.items():
        if isinstance(v, (list, tuple)):
            for i, x in enumerate(v):
                getattr(t, k)[i] = x
        else:
            setattr(t, k, v)
    return t


def cases():
    bbh = abi.source_defaults(mass1=32., mass2=21., Luminosity_Distance=450., spin1=[0, 0, .25], spin2=[0, 0, -.15], RA=1.1, DEC=-.4, psi=.7,
                              incl_angle=.9, gmst=2.1, f_ref=20., phiRef=1.3, tc=3.0, shift_time=0)
    bns = abi.source_defaults(mass1=1.6, mass2=1.3, Luminosity_Distance=80., spin1=[0, 0, .02], spin2=[0, 0, .01], RA=1.1, DEC=-.4, psi=.7,
                              incl_angle=.9, gmst=2.1, f_ref=20., phiRef=1.3, tc=3.0, shift_time=0, NSflag1=1, NSflag2=1)
    prec = _with(bbh, spin1=[.3, .1, .25], spin2=[-.1, .2, -.15])
    return [
        ("IMRPhenomD_NRT", 12, _with(bns, tidal_love=1, tidal_s=400.)),
        ("MCMC_IMRPhenomD_NRT", 12, _with(bns, tidal_love=1, tidal_s=400.)),
        ("IMRPhenomD_NRT", 13, _with(bns, tidal_love=0, tidal1=300., tidal2=500.)),
        ("gIMRPhenomD", 13, _with(bbh, Nmod_phi=1, phii=[4], delta_phi=[.05], Nmod_beta=1, betai=[2], delta_beta=[.02])),
        ("MCMC_dCS_IMRPhenomD", 12, _with(bbh, Nmod=1, bppe=[-1.], betappe=[1e-19])),
        # (EdGB_IMRPhenomD is left out on purpose: its parameter alpha^2 ~ 1e-20 s^4 is stepped by the stencil's absolute 1e-8, b = -7 turns that
        #  into phases of 1e+150 rad, and the matrix entries of that row sit at the overflow threshold in the reference itself)
        ("ppE_IMRPhenomD_IMR", 12, _with(bbh, Nmod=1, bppe=[-1.], betappe=[.01])),
        ("MCMC_ppE_IMRPhenomPv2_Inspiral", 16, _with(prec, Nmod=1, bppe=[-1.], betappe=[.01])),
    ]


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(cases())), ids=["%s_%d" % (c[0], c[1]) for c in cases()])
def test_fisher_variant_vs_oracle(ctx, oracle, k):
    method, dim, src = cases()[k]
    f = 20.0 + 0.25 * np.arange(4096)
    dets = ["Hanford", "Livingston", "Virgo"]
    psd = workloads.aligo_analytic_psd(f)[None, :] * np.array([1.0, 1.2, 2.5])[:, None]
    ctx.set_network(dets, f, psd)
    for order, di in ((4, 1), (2, 0)):
        got = ctx.fisher_numerical_batch(method, [src], dim, order=order, detector_index=di)[0]
        ref = oracle.fisher_numerical_batch(method, [src], dets, f, psd, dim, order=order, detector_index=di)[0]
        floor = fisher_noise.reference_self_difference(oracle, method, [src], dets, f, psd, dim, order, detector_index=di, runs=2)[0]
        finite = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), finite), method  # (a row the reference itself cannot form is NaN here too)
        assert finite.any(axis=1).sum() >= dim - 1, method
        rows = finite.all(axis=1)
        sub = np.ix_(rows, rows)
        err = fisher_noise.normalised_error(got[sub], ref[sub])
        assert np.median(err) <= 1e-6, (method, order, np.median(err))
        assert err.max() <= max(1e-6, fisher_noise.FACTOR * floor), (method, order, err.max(), floor)
