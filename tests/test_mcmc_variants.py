"""The sampling-vector path (MCMC_prep_params + repack_parameters("MCMC_"+method) + MCMC_likelihood_extrinsic, src/mcmc_gw.cpp:2374-2568,
src/fisher.cpp:2167-2507) for the method families and modification layouts the BASELINE configs do not reach: two tidal parameters,
several ppE terms, all four gIMR coefficient families, precessing ppE / gIMR, EdGB units, NRT with a ppE term."""
import numpy as np
import pytest

from gw_analysis_tools_b200 import abi, workloads


def variants():
    out = []
    out.append(("IMRPhenomD_NRT", 5, abi.mod_defaults(tidal_love=0, NSflag1=1, NSflag2=1), lambda r, W: np.column_stack([r.uniform(np.log(50.), np.log(2000.), W), r.uniform(np.log(50.), np.log(2000.), W)]), True))
    out.append(("ppE_IMRPhenomD_Inspiral", 1, abi.mod_defaults(ppE_Nmod=2, bppe=[-3.0, -1.0]), lambda r, W: np.column_stack([r.uniform(-1e-3, 1e-3, W), r.uniform(-0.1, 0.1, W)]), False))
    out.append(("ppE_IMRPhenomD_IMR", 1, abi.mod_defaults(ppE_Nmod=1, bppe=[-1.0]), lambda r, W: r.uniform(-0.1, 0.1, (W, 1)), False))
    out.append(("gIMRPhenomD", 1, abi.mod_defaults(gIMR_Nmod_phi=1, gIMR_phii=[4], gIMR_Nmod_sigma=1, gIMR_sigmai=[2], gIMR_Nmod_beta=1, gIMR_betai=[2],
                                                   gIMR_Nmod_alpha=1, gIMR_alphai=[3]), lambda r, W: r.uniform(-0.05, 0.05, (W, 4)), False))
    out.append(("EdGB_IMRPhenomD", 1, abi.mod_defaults(ppE_Nmod=1, bppe=[-7.0]), lambda r, W: r.uniform(0.5, 8.0, (W, 1)), False))
    out.append(("ppE_IMRPhenomPv2_Inspiral", 2, abi.mod_defaults(ppE_Nmod=1, bppe=[-1.0]), lambda r, W: r.uniform(-0.1, 0.1, (W, 1)), False))
    out.append(("gIMRPhenomPv2", 2, abi.mod_defaults(gIMR_Nmod_phi=2, gIMR_phii=[3, 6]), lambda r, W: r.uniform(-0.05, 0.05, (W, 2)), False))
    out.append(("ppE_IMRPhenomD_NRT_Inspiral", 5, abi.mod_defaults(tidal_love=1, NSflag1=1, NSflag2=1, ppE_Nmod=1, bppe=[-1.0]),
                lambda r, W: np.column_stack([r.uniform(np.log(50.), np.log(2000.), W), r.uniform(-0.05, 0.05, W)]), True))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(variants())), ids=[v[0] + "_%d" % i for i, v in enumerate(variants())])
def test_mcmc_variant_vs_oracle(ctx, oracle, k):
    method, cfg, mod, extra, replaces_tail = variants()[k]
    W = 16
    wl = workloads.make(cfg, W=W, L=4096 if cfg != 5 else 32768)
    rng = np.random.default_rng(100 + k)
    base = wl.params[:, :15 if cfg == 2 else 11]
    params = np.concatenate([base, extra(rng, W)], axis=1)
    inj = np.concatenate([wl.inj[:15 if cfg == 2 else 11], np.median(extra(rng, 64), axis=0)])
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(method, inj[None, :], wl.gmst, mod)
    src[0].tc = wl.T_segment - src[0].tc
    data = ctx.coherent_response_batch(method, src)[0]
    ref_data = oracle.coherent_response(method, src[0], wl.detectors, wl.f)
    assert np.abs(data - ref_data).max() <= 1e-10 * np.abs(ref_data).max(), method
    ctx.set_network(wl.detectors, wl.f, wl.psd, data)
    got = ctx.loglike_mcmc_batch(method, params, wl.gmst, wl.T_segment, mod)
    ref = oracle.loglike_mcmc_batch(method, mod, params, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, data)
    assert np.all(np.isfinite(ref)), method
    assert (np.abs(got - ref) / np.abs(ref)).max() <= 1e-9, (method, got, ref)
    # a vector of the wrong dimension for the layout is an error, not a guess
    from gw_analysis_tools_b200 import engine
    with pytest.raises(engine.GwatB200Error):
        ctx.loglike_mcmc_batch(method, params[:, :-1], wl.gmst, wl.T_segment, mod)
