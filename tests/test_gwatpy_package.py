"""The reference's own Python package, UNMODIFIED, on this library: gwatpy (gwatpy/gwatpy/*.py under /root/reference) is imported with its
config pointing at libgwat_b200_gwatpy.so instead of libgwat.so, and its host-side calls are compared with the compiled reference.

What this pins that tests/test_gwatpy_dropin.py cannot: the argument lists as gwatpy's ctypes code really passes them (its own argtypes,
its own array conversions, its object lifetimes), and that every module of the package imports -- gwatpy binds several symbols at import.
Needs /root/reference (this container only; skipped on the GPU box), so only the calls that need no GPU are exercised here: the GPU-side
calls of the same library are covered through the same symbols in tests/test_gwatpy_dropin.py."""
import ctypes as C
import importlib
import os
import sys
import types
from unittest import mock

import numpy as np
import pytest

from gw_analysis_tools_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gw_analysis_tools_b200", "libgwat_b200_gwatpy.so")
PKG = "/root/reference/gwatpy"


@pytest.fixture(scope="module")
def gwatpy_pkg():
    if not os.path.isdir(os.path.join(PKG, "gwatpy")):
        pytest.skip("reference not mounted")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.path.insert(0, PKG)
    cfg = types.ModuleType("gwatpy.config")  # the package's one site-specific file (gwatpy/gwatpy/config.py: LIB, GWATPY_ROOT_DIRECTORY)
    cfg.LIB = LIB
    cfg.GWATPY_ROOT_DIRECTORY = os.path.join(PKG, "gwatpy") + "/"
    pkg = importlib.import_module("gwatpy")
    sys.modules["gwatpy.config"] = cfg
    pkg.config = cfg
    for name in ("h5py", "matplotlib", "matplotlib.pyplot", "emcee", "corner"):  # plotting / file dependencies of mcmc_routines, unused here
        try:
            importlib.import_module(name)
        except ImportError:
            sys.modules[name] = mock.MagicMock()
    mods = {m: importlib.import_module("gwatpy." + m) for m in ("util", "detector_util", "waveform_generator", "mcmc_routines")}
    yield types.SimpleNamespace(**mods)
    sys.path[:] = saved_path
    for name in list(sys.modules):
        if name not in saved_mods:
            del sys.modules[name]


def _flat(obj):
    lib = C.CDLL(LIB)
    out = abi.Source()
    lib.gen_params_base_get_flat_py(C.c_void_p(obj), C.byref(out))
    return out


def test_all_modules_import_against_this_library(gwatpy_pkg):
    for mod in (gwatpy_pkg.util, gwatpy_pkg.detector_util, gwatpy_pkg.waveform_generator, gwatpy_pkg.mcmc_routines):
        assert mod.rlib._name == LIB


def test_gen_params_as_gwatpy_builds_it(gwatpy_pkg):
    """gwatpy.util.gen_params (util.py:197-398) -> gen_params_base_py: every field lands where gwatpy put it."""
    kw = dict(mass1=36.0, mass2=29.0, spin1=[0.1, 0.2, 0.3], spin2=[0.0, -0.1, -0.2], Luminosity_Distance=410.0, RA=1.2, DEC=-0.3, psi=0.4,
              incl_angle=0.5, tc=2.0, phiRef=1.0, f_ref=20.0, gmst=2.1, theta_l=0.7, phi_l=0.9, cosmology="WMAP9", NSflag1=True,
              shift_time=False, Nmod=2, bppe=[-1, 3], betappe=[0.25, -0.5], Nmod_phi=1, phii=[4], delta_phi=[0.05])
    gp = gwatpy_pkg.util.gen_params(**kw)
    s = _flat(gp.obj)
    for name in ("mass1", "mass2", "Luminosity_Distance", "RA", "DEC", "psi", "incl_angle", "tc", "phiRef", "f_ref", "gmst", "theta_l", "phi_l"):
        assert getattr(s, name) == kw[name], name
    assert list(s.spin1) == kw["spin1"] and list(s.spin2) == kw["spin2"]
    assert s.NSflag1 == 1 and s.NSflag2 == 0 and s.shift_time == 0 and s.shift_phase == 1 and s.sky_average == 0
    assert s.Nmod == 2 and list(s.bppe)[:2] == [-1.0, 3.0] and list(s.betappe)[:2] == [0.25, -0.5]
    assert s.Nmod_phi == 1 and s.phii[0] == 4 and s.delta_phi[0] == 0.05
    lib = C.CDLL(LIB)
    assert s.cosmology == lib.gwat_b200_cosmology_index(b"WMAP9") if hasattr(lib, "gwat_b200_cosmology_index") else True


def test_util_helpers_through_gwatpy(gwatpy_pkg, oracle):
    u, ref = gwatpy_pkg.util, oracle.lib()
    ref.oracle_ref_dl_from_z.restype = ref.oracle_ref_t_0pn.restype = ref.oracle_ref_f_0pn.restype = C.c_double
    ref.oracle_ref_dl_from_z.argtypes = [C.c_double, C.c_char_p]
    ref.oracle_ref_t_0pn.argtypes = ref.oracle_ref_f_0pn.argtypes = [C.c_double, C.c_double]
    for z in (1e-3, 0.2, 3.0):
        want = ref.oracle_ref_dl_from_z(z, b"PLANCK15")
        assert abs(u.DL_from_Z_py(z, "PLANCK15") - want) <= 1e-14 * want
    m1, m2 = 36.0, 29.0
    mc, eta = u.calculate_chirpmass_py(m1, m2), u.calculate_eta_py(m1, m2)
    assert abs(mc - (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2) <= 1e-14 * mc and abs(eta - m1 * m2 / (m1 + m2) ** 2) <= 1e-16
    assert abs(u.calculate_mass1_py(mc, eta) - m1) <= 1e-12 * m1 and abs(u.calculate_mass2_py(mc, eta) - m2) <= 1e-12 * m2
    v = u.calculate_chirpmass_py(np.array([m1, 10.0]), np.array([m2, 8.0]))
    assert abs(v[0] - mc) <= 1e-14 * mc and len(v) == 2
    t = u.t_0PN_py(20.0, 1.2e-4)
    assert abs(t - ref.oracle_ref_t_0pn(20.0, 1.2e-4)) <= 1e-15 * t and abs(u.f_0PN_py(t, 1.2e-4) - 20.0) <= 1e-9 * 20
    gps = 1126259462.4
    assert abs(u.gps_to_GMST_radian_py(gps) - oracle.gps_to_gmst_radian(gps)) <= 1e-9


def test_detector_helpers_through_gwatpy(gwatpy_pkg, oracle, monkeypatch):
    monkeypatch.setenv("GWAT_B200_NOISE_DIR", "/root/reference/data/noise_data/currently_supported")  # the tabulated curves' CSV files
    du, ref = gwatpy_pkg.detector_util, oracle.lib()
    sig = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(9 * C.c_double)]
    ref.oracle_ref_detector_site.argtypes = [C.c_int] + sig
    for which, name in enumerate(["Hanford", "Livingston", "Virgo", "Kagra", "Indigo", "CE", "ET1"]):
        wl, wo, wloc, wD = C.c_double(), C.c_double(), (C.c_double * 3)(), (C.c_double * 9)()
        assert ref.oracle_ref_detector_site(which, C.byref(wl), C.byref(wo), wloc, wD) == 0
        lat, lon, loc, D = du.get_detector_parameters_py(name)
        assert (lat, lon) == (wl.value, wo.value) and np.array_equal(loc, np.asarray(wloc)) and np.array_equal(D, np.asarray(wD).reshape(3, 3))
    f = np.geomspace(10.0, 2000.0, 200)
    for curve in ("aLIGO_analytic", "Hanford_O1_fitted", "AdLIGODesign"):
        got = np.asarray(du.populate_noise_py(f, curve, 48))
        want = oracle.populate_noise(f, curve)
        assert np.all(np.abs(got - want) <= 1e-13 * np.abs(want)), curve


def test_mcmc_structures_through_gwatpy(gwatpy_pkg, oracle):
    """MCMC_modification_struct_py, mcmc_data_interface_py, and MCMC_prep_params_py as gwatpy/mcmc_routines.py calls them, against the compiled
    MCMC_prep_params (repack_parameters_py runs on the GPU: tests/test_gwatpy_dropin.py)."""
    mr, u = gwatpy_pkg.mcmc_routines, gwatpy_pkg.util
    kw = dict(mass1=36.0, mass2=29.0, spin1=[0.0, 0.0, 0.3], spin2=[0.0, 0.0, -0.2], Luminosity_Distance=410.0, RA=1.2, DEC=-0.3, psi=0.4,
              incl_angle=0.5, tc=2.0, phiRef=1.0, f_ref=20.0, gmst=2.1)
    for method, dim, modkw, abikw in (("IMRPhenomD", 11, {}, {}), ("ppE_IMRPhenomD_Inspiral", 12, dict(ppE_Nmod=1, bppe=[-1]), dict(ppE_Nmod=1, bppe=[-1.0])),
                                      ("IMRPhenomPv2", 15, {}, {})):
        rng = np.random.default_rng(dim)
        x = rng.uniform(0.1, 0.9, dim)
        x[6], x[7], x[8] = np.log(400.0), np.log(25.0), 0.24
        gp = u.gen_params(**kw)
        ms = mr.MCMC_modification_struct_py(**modkw)
        got_method, temp = mr.MCMC_prep_params_py(x, gp, dim, method, ms, save_gmst=False)
        temp_ref, out_ref = oracle.mcmc_prep_params(method, abi.mod_defaults(**abikw), x, abi.source_defaults(**{k: v for k, v in kw.items()}))
        assert got_method == method and np.array_equal(temp, temp_ref)
        s = _flat(gp.obj)
        for fld in ("sky_average", "f_ref", "shift_time", "shift_phase", "NSflag1", "NSflag2", "Nmod", "gmst"):
            assert getattr(s, fld) == getattr(out_ref, fld), (method, fld)
    di = mr.mcmc_data_interface_py(min_dim=11, max_dim=13, chain_id=3, nested_model_number=0, chain_number=8, RJ_step_width=0.5, burn_phase=True)
    lib = C.CDLL(LIB)
    ints, width, burn = (C.c_int * 5)(), C.c_double(), C.c_bool()
    lib.mcmc_data_interface_get_py(C.c_void_p(di.obj), ints, C.byref(width), C.byref(burn))
    assert list(ints) == [11, 13, 3, 0, 8] and burn.value is True
    assert width.value == 8.0  # the reference stores chain_number in RJ_step_width (src/gwatpy_wrapping.cpp:260), and so does this


def test_time_domain_wrappers_through_gwatpy(gwatpy_pkg):
    """time_waveform_generator / time_response_generator (waveform_generator.py:72-170): bound at import, and for the models of this path the
    reference's answer (nothing computed) comes back."""
    wg, u = gwatpy_pkg.waveform_generator, gwatpy_pkg.util
    gp = u.gen_params(mass1=36.0, mass2=29.0)
    t = np.linspace(-1.0, 0.0, 32)
    out = wg.time_waveform_generator(t, "IMRPhenomD", gp)
    assert all(not np.any(np.asarray(a)) for a in out)
    resp = wg.time_response_generator(t, "Hanford", "IMRPhenomD", gp)
    assert not np.any(np.asarray(resp))
