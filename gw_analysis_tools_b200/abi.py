"""ctypes mirror of ``include/gwat_b200.h`` (PODs and constants only, no logic).

The field order here must match the C header exactly; ``tests/test_abi.py`` checks ``sizeof`` against the library.
Field names follow the reference's ``gen_params_base<double>`` (include/gwat/util.h:121-378) and
``MCMC_modification_struct`` (include/gwat/mcmc_gw.h:50-88).
"""
import ctypes as C

ABI_VERSION = 1
MAX_DETECTORS = 8
MAX_MOD = 8
MAX_DIM = 32

OK, ERR_ARG, ERR_METHOD, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

_dM = C.c_double * MAX_MOD
_iM = C.c_int * MAX_MOD
_d3 = C.c_double * 3


class Source(C.Structure):
    """``gwat_b200_source``: flattened ``gen_params_base<double>``."""

    _fields_ = [
        ("mass1", C.c_double), ("mass2", C.c_double),
        ("Luminosity_Distance", C.c_double),
        ("spin1", _d3), ("spin2", _d3),
        ("tc", C.c_double), ("phiRef", C.c_double), ("f_ref", C.c_double),
        ("psi", C.c_double), ("incl_angle", C.c_double),
        ("RA", C.c_double), ("DEC", C.c_double), ("gmst", C.c_double),
        ("theta", C.c_double), ("phi", C.c_double),
        ("theta_l", C.c_double), ("phi_l", C.c_double),
        ("tidal1", C.c_double), ("tidal2", C.c_double), ("tidal_s", C.c_double), ("tidal_a", C.c_double),
        ("tidal_weighted", C.c_double), ("delta_tidal_weighted", C.c_double),
        ("diss_tidal1", C.c_double), ("diss_tidal2", C.c_double), ("diss_tidal_s", C.c_double),
        ("diss_tidal_a", C.c_double), ("diss_tidal_weighted", C.c_double),
        ("chip", C.c_double), ("phip", C.c_double),
        ("betappe", _dM), ("bppe", _dM),
        ("delta_phi", _dM), ("delta_sigma", _dM), ("delta_beta", _dM), ("delta_alpha", _dM),
        ("phii", _iM), ("sigmai", _iM), ("betai", _iM), ("alphai", _iM),
        ("Nmod", C.c_int), ("Nmod_phi", C.c_int), ("Nmod_sigma", C.c_int), ("Nmod_beta", C.c_int),
        ("Nmod_alpha", C.c_int),
        ("PNorder", C.c_int),
        ("shift_time", C.c_int), ("shift_phase", C.c_int), ("sky_average", C.c_int),
        ("tidal_love", C.c_int), ("tidal_love_error", C.c_int),
        ("NSflag1", C.c_int), ("NSflag2", C.c_int),
        ("dep_postmerger", C.c_int),
        ("equatorial_orientation", C.c_int), ("horizon_coord", C.c_int),
        ("cosmology", C.c_int),
        ("reserved_", C.c_int),
    ]


class Mod(C.Structure):
    """``gwat_b200_mod``: flattened ``MCMC_modification_struct``."""

    _fields_ = [
        ("ppE_Nmod", C.c_int),
        ("bppe", _dM),
        ("gIMR_Nmod_phi", C.c_int), ("gIMR_Nmod_sigma", C.c_int), ("gIMR_Nmod_beta", C.c_int),
        ("gIMR_Nmod_alpha", C.c_int),
        ("gIMR_phii", _iM), ("gIMR_sigmai", _iM), ("gIMR_betai", _iM), ("gIMR_alphai", _iM),
        ("NSflag1", C.c_int), ("NSflag2", C.c_int),
        ("tidal_love", C.c_int), ("tidal_love_error", C.c_int),
    ]


def source_defaults(**kw):
    """A ``Source`` with the reference's member defaults (include/gwat/util.h:125-285), then ``kw`` applied."""
    s = Source()
    s.tidal1 = s.tidal2 = s.tidal_s = s.tidal_a = s.tidal_weighted = s.delta_tidal_weighted = -1.0
    s.diss_tidal1 = s.diss_tidal2 = s.diss_tidal_s = s.diss_tidal_a = s.diss_tidal_weighted = -1.0
    s.chip = -1.0
    s.phip = -1.0
    s.PNorder = 35
    s.shift_time = 1
    s.shift_phase = 1
    s.tidal_love = 1
    for k, v in kw.items():
        cur = getattr(s, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(s, k, v)
    return s


def mod_defaults(**kw):
    """A ``Mod`` with the reference's defaults (include/gwat/mcmc_gw.h:52-67), then ``kw`` applied."""
    m = Mod()
    m.tidal_love = 1
    for k, v in kw.items():
        cur = getattr(m, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(m, k, v)
    return m
