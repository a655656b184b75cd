#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the ORACLE (the reference's own code, oracle/_ref/libgwat_ref.so).

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py
The vectors pin parity on boxes where the oracle library is absent, and pin the oracle itself against regressions of the
stub headers / build recipe.  Everything is double precision; nothing is rounded on the way to disk.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from gw_analysis_tools_b200 import workloads  # noqa: E402
from oracle import gwat_ref as R  # noqa: E402


def maximized():
    """tc/phic-maximised likelihoods of every family case against the case's derived data."""
    gold = np.load(os.path.join(HERE, "waveforms_v1.npz"))
    mx = {}
    for name, method, kw, gspec in cases.CASES:
        f = cases.grid(gspec)
        src = cases.source_from_bytes(gold[name + "/src"])
        data = cases.maximized_data(gold, gspec)
        psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
        mx[name] = R.loglike_maximized_batch(method, [src], cases.DETECTORS, f, psd, data)[0]
        print("%-12s %-28s maximised logL %.12e" % (name, method, mx[name]))
    np.savez_compressed(os.path.join(HERE, "maximized_v1.npz"), **mx)


def theories():
    """The theory mappings beyond dCS/EdGB: polarisations and the likelihood against derived data."""
    out = {}
    gold = np.load(os.path.join(HERE, "waveforms_v1.npz"))
    for name, method, kw, gspec in cases.THEORY_CASES:
        f = cases.grid(gspec)
        src = cases.source(kw)
        hp, hc = R.fourier_waveform(method, src, f)
        psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
        # data of a fixed GR case per grid: the template is mismatched, so the likelihood feels the theory's phase
        ll = R.loglike_batch(method, [src], cases.DETECTORS, f, psd, cases.maximized_data(gold, gspec))[0]
        gr = R.fourier_waveform(_gr_method(method), cases.source({k: v for k, v in kw.items() if k not in ("Nmod", "bppe", "betappe")}), f)[0]
        out[name + "/src"] = cases.source_bytes(src)
        out[name + "/hp"] = hp
        out[name + "/hc"] = hc
        out[name + "/logL"] = np.array(ll)
        dphi = np.abs(np.angle(hp[np.abs(gr) > 0] / gr[np.abs(gr) > 0])).max()
        print("%-14s %-34s logL %.12e   max dephasing vs GR %.3g rad" % (name, method, ll, dphi))
    np.savez_compressed(os.path.join(HERE, "theories_v1.npz"), **out)


def _gr_method(method):
    for base in ("IMRPhenomD_NRT", "IMRPhenomPv2", "IMRPhenomD"):
        if base in method:
            return base
    raise ValueError(method)


def fisher_noise():
    """The self-difference yardstick of every Fisher golden matrix (tests/fisher_noise.py): largest normalised difference among
    the FMA-contracted build and four re-evaluations on inputs moved by parts in 1e14."""
    import fisher_noise as FN
    f = cases.grid(cases.FISHER_GRID)
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    out = {}
    for name, method, kw, dim in cases.FISHER_CASES:
        src = cases.source(kw)
        row = []
        for order in (2, 4):
            for d, det in list(enumerate(cases.DETECTORS[:2])) + [(-1, "sum")]:
                if d < 0 and order == 2:
                    continue
                v = FN.reference_self_difference(R, method, [src], cases.DETECTORS, f, psd, dim, order, detector_index=d, nthreads=1)[0]
                out["%s/o%d/%s" % (name, order, det)] = np.array(v)
                row.append("%.1e" % v)
        print("%-10s self-difference %s" % (name, " ".join(row)))
    np.savez_compressed(os.path.join(HERE, "fisher_noise_v2.npz"), **out)


def main():
    if "--only-fisher-noise" in sys.argv:
        return fisher_noise()
    if "--only-maximized" in sys.argv:
        return maximized()
    if "--only-theories" in sys.argv:
        return theories()
    out = {}
    for name, method, kw, gspec in cases.CASES:
        f = cases.grid(gspec)
        src = cases.source(kw)
        hp, hc = R.fourier_waveform(method, src, f)
        resp = R.coherent_response(method, src, cases.DETECTORS, f)
        data = cases.derived_data(resp)
        psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
        ll = R.loglike_batch(method, [src], cases.DETECTORS, f, psd, data)[0]
        out[name + "/src"] = cases.source_bytes(src)
        out[name + "/hp"] = hp
        out[name + "/hc"] = hc
        out[name + "/resp"] = resp
        if name in ("D_bbh", "P_full", "NRT_love"):
            out[name + "/single_L"] = R.fourier_detector_response(method, "Livingston", src, f)
        out[name + "/logL"] = np.array(ll)
        print("%-12s %-28s max|h+| %.3e  logL %.12e" % (name, method, np.abs(hp).max(), ll))
    np.savez_compressed(os.path.join(HERE, "waveforms_v1.npz"), **out)

    fo = {}
    f = cases.grid(cases.FISHER_GRID)
    psd = np.tile(workloads.aligo_analytic_psd(f), (3, 1))
    for name, method, kw, dim in cases.FISHER_CASES:
        src = cases.source(kw)
        for order in (2, 4):
            for d, det in enumerate(cases.DETECTORS[:2]):
                F = R.fisher_numerical_batch(method, [src], cases.DETECTORS, f, psd, dim, order=order, detector_index=d,
                                             reference_index=0, nthreads=1)[0]
                fo["%s/o%d/%s" % (name, order, det)] = F
                # the reference's own rounding-noise floor: the same sources rebuilt with FMA contraction
                Ff = R.fisher_numerical_batch(method, [src], cases.DETECTORS, f, psd, dim, order=order, detector_index=d,
                                              reference_index=0, nthreads=1, fma_build=True)[0]
                dg = np.sqrt(np.abs(np.diag(F)))
                fo["%s/o%d/%s/noise" % (name, order, det)] = np.array((np.abs(Ff - F) / np.outer(dg, dg)).max())
        fo[name + "/sum_o4"] = R.fisher_numerical_batch(method, [src], cases.DETECTORS, f, psd, dim, order=4, detector_index=-1,
                                                        reference_index=0, nthreads=1)[0]
        fo[name + "/src"] = cases.source_bytes(src)
        print("%-10s %-28s F00 %.6e  noise floors %s" % (name, method, fo[name + "/sum_o4"][0, 0], " ".join(
            "%.1e" % fo["%s/o%d/%s/noise" % (name, o, dt)] for o in (2, 4) for dt in cases.DETECTORS[:2])))
    np.savez_compressed(os.path.join(HERE, "fisher_v1.npz"), **fo)

    # MCMC-parameterised likelihood batches (the bench's call path) for small versions of the BASELINE configs
    mo = {}
    for cfg in (1, 2, 4, 5):
        L = 1024 if cfg != 5 else 4096
        wl = workloads.make(cfg, W=16, L=L)
        _, src = R.loglike_mcmc_batch(wl.method, wl.mod, wl.inj[None, :], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, None,
                                      return_sources=True)
        data = R.coherent_response(wl.method, src[0], wl.detectors, wl.f)
        ll = R.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, data)
        mo["cfg%d/data" % cfg] = data
        mo["cfg%d/logL" % cfg] = ll
        mo["cfg%d/params" % cfg] = wl.params
        print("cfg%d %-18s logL[0:3] %s" % (cfg, wl.method, ll[:3]))
    # the smoke() workload
    wl = workloads.make(2, W=32, L=2048)
    _, src = R.loglike_mcmc_batch(wl.method, wl.mod, wl.inj[None, :], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, None,
                                  return_sources=True)
    data = R.coherent_response(wl.method, src[0], wl.detectors, wl.f)
    ll = R.loglike_mcmc_batch(wl.method, wl.mod, wl.params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, data)
    np.savez_compressed(os.path.join(HERE, "smoke_cfg2.npz"), logL=ll, data=data)
    np.savez_compressed(os.path.join(HERE, "mcmc_v1.npz"), **mo)
    maximized()
    theories()
    fisher_noise()
    for fn in ("fisher_noise_v2.npz", "theories_v1.npz", "waveforms_v1.npz", "fisher_v1.npz", "mcmc_v1.npz", "smoke_cfg2.npz", "maximized_v1.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
