"""Shared test cases: one physical source per waveform family, used by the golden-vector generator and the parity tests."""
import ctypes as C

import numpy as np

from gw_analysis_tools_b200 import abi

BBH = dict(mass1=36.4, mass2=29.3, Luminosity_Distance=500.0, spin1=[0, 0, .3], spin2=[0, 0, .2], RA=.275, DEC=-.44, psi=.2,
           incl_angle=.51, gmst=2.1, f_ref=20.0, phiRef=2.0, tc=3.0)
BBH_LOW = dict(BBH, mass1=10.0, mass2=8.0, Luminosity_Distance=300.0, spin1=[0, 0, -.4], spin2=[0, 0, .6], tc=5.5)
BBH_Q8 = dict(BBH, mass1=64.0, mass2=8.0, spin1=[0, 0, .85], spin2=[0, 0, -.5])
PREC = dict(BBH, spin1=[.3, .1, .3], spin2=[0, -.2, .2])
BNS = dict(mass1=1.5, mass2=1.3, Luminosity_Distance=100.0, spin1=[0, 0, .02], spin2=[0, 0, .01], RA=1.1, DEC=.3, psi=1.2,
           incl_angle=2.4, gmst=2.1, f_ref=20.0, phiRef=.7, tc=6.0)
KM4 = lambda km: (km / 299792.458) ** 4  # sqrt(alpha) in km -> alpha^2 in s^4

# name, method, source kwargs, (fmin, df, L)
GRID_BBH = (20.0, 2.0, 512)
GRID_BNS = (10.0, 2.0, 1024)
CASES = [
    ("D_bbh", "IMRPhenomD", BBH, GRID_BBH),
    ("D_low", "IMRPhenomD", BBH_LOW, GRID_BBH),
    ("D_q8", "IMRPhenomD", BBH_Q8, GRID_BBH),
    ("D_noshift", "IMRPhenomD", dict(BBH, shift_time=0, shift_phase=0), GRID_BBH),
    ("P_full", "IMRPhenomPv2", PREC, GRID_BBH),
    ("P_reduced", "IMRPhenomPv2", dict(BBH, chip=0.4, phip=1.1), GRID_BBH),
    ("P_low", "IMRPhenomPv2", dict(BBH_LOW, spin1=[-.2, .5, -.4], spin2=[.1, .1, .6]), GRID_BBH),
    ("P_nrt", "IMRPhenomPv2_NRT", PREC, GRID_BBH),
    ("ppE_ins", "ppE_IMRPhenomD_Inspiral", dict(BBH, Nmod=2, bppe=[-7., -3.], betappe=[1e-6, 0.02]), GRID_BBH),
    ("ppE_imr", "ppE_IMRPhenomD_IMR", dict(BBH, Nmod=2, bppe=[1., 3.], betappe=[0.5, -2.]), GRID_BBH),
    ("gIMR", "gIMRPhenomD", dict(BBH, Nmod_phi=3, phii=[-2, 1, 6], delta_phi=[0.01, 0.05, 0.1], Nmod_sigma=1, sigmai=[2],
                                 delta_sigma=[0.1], Nmod_beta=1, betai=[3], delta_beta=[-.1], Nmod_alpha=2, alphai=[2, 4],
                                 delta_alpha=[.1, .2]), GRID_BBH),
    ("gIMR_log", "gIMRPhenomD", dict(BBH, Nmod_phi=2, phii=[8, 9], delta_phi=[0.3, -0.2]), GRID_BBH),
    ("ppE_P_ins", "ppE_IMRPhenomPv2_Inspiral", dict(PREC, Nmod=1, bppe=[-1.], betappe=[0.1]), GRID_BBH),
    ("ppE_P_imr", "ppE_IMRPhenomPv2_IMR", dict(PREC, Nmod=1, bppe=[-1.], betappe=[0.1]), GRID_BBH),
    ("gIMR_P", "gIMRPhenomPv2", dict(PREC, Nmod_phi=2, phii=[-1, 3], delta_phi=[0.05, .1], Nmod_alpha=1, alphai=[3],
                                     delta_alpha=[.1]), GRID_BBH),
    ("dCS", "dCS_IMRPhenomD", dict(BBH, Nmod=1, bppe=[-1.], betappe=[KM4(20.)]), GRID_BBH),
    ("EdGB", "EdGB_IMRPhenomD", dict(BBH_LOW, Nmod=1, bppe=[-7.], betappe=[KM4(3.)]), GRID_BBH),
    ("dCS_P", "dCS_IMRPhenomPv2", dict(PREC, Nmod=1, bppe=[-1.], betappe=[KM4(20.)]), GRID_BBH),
    ("NRT_love", "IMRPhenomD_NRT", dict(BNS, tidal_love=1, tidal_s=400.), GRID_BNS),
    ("NRT_12", "IMRPhenomD_NRT", dict(BNS, tidal_love=0, tidal1=300., tidal2=500., diss_tidal1=10., diss_tidal2=20.), GRID_BNS),
    ("NRT_w", "IMRPhenomD_NRT", dict(BNS, tidal_love=0, tidal_weighted=350.), GRID_BNS),
    ("ppE_NRT_ins", "ppE_IMRPhenomD_NRT_Inspiral", dict(BNS, tidal_love=1, tidal_s=400., Nmod=1, bppe=[-1.], betappe=[0.01]), GRID_BNS),
    ("ppE_NRT_imr", "ppE_IMRPhenomD_NRT_IMR", dict(BNS, tidal_love=1, tidal_s=400., Nmod=1, bppe=[-1.], betappe=[0.01]), GRID_BNS),
]
# The other theory mappings of assign_mapping (src/ppE_utilities.cpp:158-359); couplings chosen for O(0.1-1) rad of dephasing.
THEORY_CASES = [
    ("EdGB_HO", "EdGB_HO_IMRPhenomD", dict(BBH_LOW, Nmod=1, bppe=[-7.], betappe=[KM4(3.)]), GRID_BBH),  # == EdGB (reference quirk)
    ("EdGB_HO_LO", "EdGB_HO_LO_IMRPhenomD", dict(BBH_LOW, Nmod=1, bppe=[-7.], betappe=[KM4(2.)]), GRID_BBH),
    ("EdGB_GHOv1", "EdGB_GHOv1_IMRPhenomD", dict(BBH_LOW, Nmod=2, bppe=[-7., -5.], betappe=[KM4(3.), -2e-4]), GRID_BBH),
    ("EdGB_GHOv2", "EdGB_GHOv2_IMRPhenomD", dict(BBH_LOW, Nmod=2, bppe=[-7., -5.], betappe=[KM4(3.), 2.0]), GRID_BBH),
    ("EdGB_GHOv3", "EdGB_GHOv3_IMRPhenomPv2", dict(PREC, Nmod=2, bppe=[-7., -5.], betappe=[KM4(8.), 1.5]), GRID_BBH),
    ("ExtraDim", "ExtraDimension_IMRPhenomD", dict(BBH, Nmod=1, bppe=[-13.], betappe=[1e-10]), GRID_BBH),
    ("BHEvap", "BHEvaporation_IMRPhenomD", dict(BBH, Nmod=1, bppe=[-13.], betappe=[2e-6]), GRID_BBH),
    ("TVG", "TVG_IMRPhenomPv2", dict(PREC, Nmod=1, bppe=[-13.], betappe=[1e-2]), GRID_BBH),
    ("DipRad", "DipRad_IMRPhenomD", dict(BBH, Nmod=1, bppe=[-7.], betappe=[1e-3]), GRID_BBH),
    ("DipRad_NRT", "DipRad_IMRPhenomD_NRT", dict(BNS, tidal_love=1, tidal_s=400., Nmod=1, bppe=[-7.], betappe=[1e-5]), GRID_BNS),
    ("NonComm", "NonComm_IMRPhenomD", dict(BBH, Nmod=1, bppe=[-1.], betappe=[0.5]), GRID_BBH),
    ("PNSeries_ins", "PNSeries_ppE_IMRPhenomD_Inspiral", dict(BBH, Nmod=2, bppe=[-7., -5.], betappe=[1e-6, 50.]), GRID_BBH),
    ("PNSeries_imr", "PNSeries_ppE_IMRPhenomPv2_IMR", dict(PREC, Nmod=2, bppe=[-1., 1.], betappe=[0.1, 0.5]), GRID_BBH),
    ("ppEAlt_ins", "ppEAlt_IMRPhenomD_Inspiral", dict(BBH, Nmod=2, bppe=[-7., -3.], betappe=[1e-6, 0.02]), GRID_BBH),
    ("ppEAlt_imr", "ppEAlt_IMRPhenomD_IMR", dict(BBH, Nmod=2, bppe=[-1., 1.], betappe=[0.1, 0.5]), GRID_BBH),
    # modified dispersion (arXiv:1110.2720): b = 3 alpha - 3 selects the table of D_alpha(z); betappe[0] = A_alpha
    ("ModDisp_a0", "ModDispersion_IMRPhenomD", dict(BBH, Nmod=1, bppe=[-3.], betappe=[1e-45]), GRID_BBH),
    ("ModDisp_a05", "ModDispersion_IMRPhenomD", dict(BBH_LOW, Nmod=1, bppe=[-1.5], betappe=[1e-39]), GRID_BBH),
    ("ModDisp_a15", "ModDispersion_IMRPhenomPv2", dict(PREC, Nmod=1, bppe=[1.5], betappe=[1e-26]), GRID_BBH),
    ("ModDisp_a2", "ModDispersion_IMRPhenomD", dict(BBH, Nmod=1, bppe=[3.], betappe=[5e-21]), GRID_BBH),  # linear in f: absorbed by the time shift
]
DETECTORS = ["Hanford", "Livingston", "Virgo"]

# Fisher cases: name, method string as the reference takes it, source kwargs, dimension
FISHER_CASES = [
    ("F_D", "IMRPhenomD", BBH, 11),
    ("F_D_mcmc", "MCMC_IMRPhenomD", BBH_LOW, 11),
    ("F_P_mcmc", "MCMC_IMRPhenomPv2", PREC, 15),
    ("F_P_red", "IMRPhenomPv2", dict(BBH, chip=0.4, phip=1.1), 13),
    ("F_ppE", "ppE_IMRPhenomD_Inspiral", dict(BBH, Nmod=1, bppe=[-1.], betappe=[0.1]), 12),
]
FISHER_GRID = (20.0, 0.25, 4096)


def derived_data(resp):
    """Deterministic 'data' for the likelihood cases, derived from the stored responses (keeps the fixtures small)."""
    return 0.9 * np.exp(0.3j) * resp


def maximized_data(gold, gspec):
    """Data for the maximised-likelihood cases: the response of ONE fixed case per grid (plain IMRPhenomD / IMRPhenomD_NRT),
    so that every other family is a mismatched template and its phase modifications matter."""
    return derived_data(gold[("D_bbh" if tuple(gspec) == tuple(GRID_BBH) else "NRT_love") + "/resp"])


def grid(spec):
    fmin, df, L = spec
    return fmin + df * np.arange(L, dtype=np.float64)


def source(kw):
    return abi.source_defaults(**kw)


def source_bytes(src):
    return np.frombuffer(C.string_at(C.addressof(src), C.sizeof(src)), dtype=np.uint8).copy()


def source_from_bytes(b):
    s = abi.Source()
    C.memmove(C.addressof(s), np.ascontiguousarray(b).ctypes.data, C.sizeof(s))
    return s
