"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck) runs on a GPU box:
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gw_analysis_tools_b200 import abi, engine, sampler, workloads  # noqa: E402


def main():
    ctx = engine.Context(0)
    for cfg in (1, 2, 4, 5):
        wl = workloads.make(cfg, W=48, L=3000 if cfg != 5 else 70000)  # odd sizes: ragged tiles and a partial last unit
        ctx.set_network(wl.detectors, wl.f, wl.psd)
        src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
        src[0].tc = wl.T_segment - src[0].tc
        data = ctx.coherent_response_batch(wl.method, src)[0]
        ctx.set_network(wl.detectors, wl.f, wl.psd, data)
        ll = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
        print("cfg%d logL finite: %d of %d" % (cfg, np.isfinite(ll).sum(), ll.size), flush=True)
        ctx.fourier_waveform_batch(wl.method, list(src))
        ctx.snr_batch(wl.method, list(src))
    # Fisher: fused kernel (pointed, orientation, intrinsic Pv2), two-kernel path (sky-averaged), maximised likelihood
    f = 20 + 0.25 * np.arange(1500)
    psd = np.tile(workloads.aligo_analytic_psd(f), (2, 1))
    ctx.set_network(["Hanford", "Virgo"], f, psd)
    srcs = workloads.fisher_sources(9)
    F = ctx.fisher_numerical_batch("IMRPhenomD", srcs, 11, order=4, detector_index=-1)
    print("Fisher finite:", np.isfinite(F).all(), flush=True)
    ctx.fisher_numerical_batch("IMRPhenomD", srcs, 11, order=2, detector_index=1)
    eq = list(srcs)
    for s in eq:
        s.equatorial_orientation, s.theta_l, s.phi_l = 1, 1.0, 2.0
    ctx.fisher_numerical_batch("MCMC_IMRPhenomD", eq, 11, order=4, detector_index=-1)
    pars = np.array([[np.log(20.), 0.22, 0.5, 0.4, 0.3, -0.5, 1.0, 4.0]] * 5)
    ctx.fisher_numerical_batch("MCMC_IMRPhenomPv2", list(ctx.repack_mcmc_intrinsic_batch("IMRPhenomPv2", pars, 2.1)), 8, order=4, detector_index=-1)
    sky = [abi.source_defaults(mass1=30., mass2=20., Luminosity_Distance=400., spin1=[0, 0, .1], spin2=[0, 0, .2], f_ref=20., sky_average=1)] * 3
    ctx.fisher_numerical_batch("IMRPhenomD", sky, 7, order=4, detector_index=0)
    ctx.fisher_numerical_batch("MCMC_IMRPhenomD", sky, 4, order=4, detector_index=0)
    inj = abi.source_defaults(mass1=31., mass2=24., Luminosity_Distance=400., RA=1., DEC=.3, psi=.4, incl_angle=.6, gmst=2.1, f_ref=20., phiRef=1.3, tc=2.5)
    ctx.set_network(["Hanford", "Virgo"], f, psd, ctx.coherent_response_batch("IMRPhenomD", [inj])[0])
    print("maximised:", ctx.loglike_maximized_mcmc_batch("IMRPhenomD", np.array([[np.log(23.), 0.24, 0.1, 0.0]] * 4), 2.1), flush=True)
    F4 = sampler.mcmc_fisher_intrinsic_batch(ctx, "IMRPhenomD", np.array([[np.log(23.), 0.24, 0.1, 0.0]] * 2), 2.1)
    print("intrinsic Fisher finite:", np.isfinite(F4).all(), flush=True)
    # the device sampler: a few steps with swaps and Fisher refreshes
    wl = workloads.make(1, W=64, L=2048)
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    ctx.set_network(wl.detectors, wl.f, wl.psd, ctx.coherent_response_batch(wl.method, src)[0])
    temps = np.tile(np.geomspace(1.0, 30.0, 8), 8)
    s = sampler.Sampler(ctx, wl.method, temps, wl.params[:64], sampler.prior_for(wl), wl.gmst, wl.T_segment, wl.mod, seed=5, fisher_exist=1,
                        swp_freq=3, history_length=12, history_update=2, fisher_update_number=4, check_stepsize_freq=5)
    s.run(12)
    print("sampler ran", flush=True)
    # autocorrelation lengths: ragged length (L = 2048 for 1025 steps), a trim, a row count that is not a multiple of anything
    rng = np.random.default_rng(2)
    ac, tau = sampler.autocorrelation_lengths(ctx, np.cumsum(rng.standard_normal((3, 1030, 5)), axis=1), begin=5)
    print("autocorrelation lags:", ac.min(), ac.max(), flush=True)
    ctx.close()
    print("sanitize_small done", flush=True)


if __name__ == "__main__":
    main()
