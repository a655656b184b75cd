"""TEST INFRASTRUCTURE ONLY: ctypes binding of the oracle ``oracle/_ref/libgwat_ref.so``.

That library is the reference's own C++ sources (compiled from /root/reference by ``oracle/Makefile``) behind the small
C driver ``oracle/ref_driver.cpp``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; nothing under ``gw_analysis_tools_b200/`` does.
"""
import ctypes as C
import os

import numpy as np

from gw_analysis_tools_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgwat_ref.so")

LIB_PATH_FMA = os.path.join(_HERE, "_ref", "libgwat_ref_fma.so")  # `make -C oracle noise`: same sources, FMA contraction on

_dp = C.POINTER(C.c_double)
_lib = None
_lib_fma = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle not built: run `make -C oracle` (needs /root/reference) -> " + LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_ref_log_likelihood_internal.restype = C.c_double
        _lib.oracle_ref_sizeof_source.restype = C.c_size_t
        _lib.oracle_ref_sizeof_mod.restype = C.c_size_t
        assert _lib.oracle_ref_sizeof_source() == C.sizeof(abi.Source)
        assert _lib.oracle_ref_sizeof_mod() == C.sizeof(abi.Mod)
    return _lib


def lib_fma():
    """The FMA-contracted build of the same reference sources: used only to measure the reference's own rounding noise."""
    global _lib_fma
    if _lib_fma is None:
        if not os.path.exists(LIB_PATH_FMA):
            raise RuntimeError("noise-floor oracle not built: run `make -C oracle noise`")
        _lib_fma = C.CDLL(LIB_PATH_FMA)
    return _lib_fma


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _dets(names):
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    return arr


def _src_array(sources):
    if isinstance(sources, abi.Source):
        sources = [sources]
    if isinstance(sources, C.Array):
        return sources, len(sources)
    arr = (abi.Source * len(sources))(*sources)
    return arr, len(sources)


def max_threads():
    return lib().oracle_ref_max_threads()


def fourier_waveform(method, src, f):
    f = _f64(f)
    L = f.size
    out = [np.zeros(L) for _ in range(4)]
    lib().oracle_ref_fourier_waveform(method.encode(), C.byref(src), _p(f), L, *[_p(o) for o in out])
    return out[0] + 1j * out[1], out[2] + 1j * out[3]


def fourier_detector_response(method, detector, src, f):
    f = _f64(f)
    L = f.size
    re, im = np.zeros(L), np.zeros(L)
    lib().oracle_ref_fourier_detector_response(method.encode(), detector.encode(), C.byref(src), _p(f), L, _p(re), _p(im))
    return re + 1j * im


def coherent_response(method, src, detectors, f):
    f = _f64(f)
    L, D = f.size, len(detectors)
    re, im = np.zeros((D, L)), np.zeros((D, L))
    lib().oracle_ref_coherent_response(method.encode(), C.byref(src), D, _dets(detectors), _p(f), L, _p(re), _p(im))
    return re + 1j * im


def log_likelihood_internal(data, psd, f, weights, resp, log10F=False, integ="SIMPSONS"):
    f = _f64(f)
    dre, dim = _f64(data.real), _f64(data.imag)
    rre, rim = _f64(resp.real), _f64(resp.imag)
    psd, weights = _f64(psd), _f64(weights)
    return lib().oracle_ref_log_likelihood_internal(_p(dre), _p(dim), _p(psd), _p(f), _p(weights), _p(rre), _p(rim),
                                                     f.size, int(log10F), integ.encode())


def loglike_batch(method, sources, detectors, f, psd, data, weights=None, integ="SIMPSONS", log10F=False, nthreads=0):
    arr, W = _src_array(sources)
    f, psd, weights = _f64(f), _f64(psd), _f64(weights)
    dre, dim = _f64(data.real), _f64(data.imag)
    out = np.zeros(W)
    lib().oracle_ref_loglike_batch(method.encode(), W, arr, len(detectors), _dets(detectors), _p(f), f.size, _p(psd),
                                   _p(dre), _p(dim), _p(weights), integ.encode(), int(log10F), int(nthreads), _p(out))
    return out


def loglike_mcmc_batch(method, mod, params, gmst, T_segment, detectors, f, psd, data, weights=None, integ="SIMPSONS",
                       log10F=False, nthreads=0, return_sources=False):
    params = _f64(params)
    W, P = params.shape
    f, psd, weights = _f64(f), _f64(psd), _f64(weights)
    dre = _f64(data.real) if data is not None else None
    dim = _f64(data.imag) if data is not None else None
    out = np.zeros(W)
    srcs = (abi.Source * W)() if return_sources else None
    lib().oracle_ref_loglike_mcmc_batch(method.encode(), C.byref(mod) if mod is not None else None, P, W, _p(params),
                                        C.c_double(gmst), C.c_double(T_segment), len(detectors), _dets(detectors),
                                        _p(f), f.size, _p(psd), _p(dre), _p(dim), _p(weights), integ.encode(),
                                        int(log10F), int(nthreads), _p(out) if data is not None else None, srcs)
    return (out, srcs) if return_sources else out


def loglike_maximized_batch(method, sources, detectors, f, psd, data, nthreads=0):
    """The reference's tc/phic-maximised likelihoods (src/mcmc_gw.cpp:595-795) as its intrinsic samplers call them."""
    arr, W = _src_array(sources)
    f, psd = _f64(f), _f64(psd)
    dre, dim = _f64(data.real), _f64(data.imag)
    out = np.zeros(W)
    lib().oracle_ref_loglike_maximized_batch(method.encode(), W, arr, len(detectors), _dets(detectors), _p(f), f.size, _p(psd),
                                             _p(dre), _p(dim), int(nthreads), _p(out))
    return out


def fisher_numerical_batch(method, sources, detectors, f, psd, dimension, order=4, detector_index=-1,
                           reference_index=0, nthreads=0, fma_build=False):
    arr, S = _src_array(sources)
    f, psd = _f64(f), _f64(psd)
    out = np.zeros((S, dimension, dimension))
    (lib_fma() if fma_build else lib()).oracle_ref_fisher_numerical_batch(method.encode(), detector_index, reference_index, dimension, order, S, arr,
                                            len(detectors), _dets(detectors), _p(f), f.size, _p(psd), int(nthreads),
                                            _p(out))
    return out


def antenna_batch(RA, DEC, psi, gmst, detectors):
    RA, DEC, psi = _f64(RA), _f64(DEC), _f64(psi)
    W, D = RA.size, len(detectors)
    fp, fc, dt = np.zeros((W, D)), np.zeros((W, D)), np.zeros((W, D))
    lib().oracle_ref_antenna_batch(W, _p(RA), _p(DEC), _p(psi), C.c_double(gmst), D, _dets(detectors), _p(fp), _p(fc),
                                   _p(dt))
    return fp, fc, dt


def gauleg_grid(f_lower, f_upper, n, log10F=True):
    f, w = np.zeros(n), np.zeros(n)
    lib().oracle_ref_gauleg_grid(C.c_double(f_lower), C.c_double(f_upper), int(n), int(bool(log10F)), _p(f), _p(w))
    return f, w


def populate_noise(f, curve):
    f = _f64(f)
    asd = np.zeros(f.size)
    lib().oracle_ref_populate_noise(_p(f), curve.encode(), _p(asd), f.size)
    return asd


def fourier_amplitude_phase(method, src, f):
    f = _f64(f)
    a, p = np.zeros(f.size), np.zeros(f.size)
    lib().oracle_ref_fourier_amplitude_phase(method.encode(), C.byref(src), _p(f), f.size, _p(a), _p(p))
    return a, p


def gps_to_gmst_radian(gps):
    fn = lib().oracle_ref_gps_to_gmst_radian
    fn.restype = C.c_double
    fn.argtypes = [C.c_double]
    return fn(float(gps))


def losc(data_files, psd_file, trigger_time, post_merger_duration, psd_length, data_file_length):
    """allocate_LOSC_data: (frequencies[L], psd[D][L], data[D][L])."""
    D = len(data_files)
    names = (C.c_char_p * D)(*[str(p).encode() for p in data_files])
    f, psd, dre, dim = np.zeros(psd_length), np.zeros((D, psd_length)), np.zeros((D, psd_length)), np.zeros((D, psd_length))
    lib().oracle_ref_losc(D, names, str(psd_file).encode(), C.c_double(trigger_time), C.c_double(post_merger_duration), int(psd_length),
                          int(data_file_length), _p(f), _p(psd), _p(dre), _p(dim))
    return f, psd, dre + 1j * dim


def calculate_snr(curve, detector, method, src, f, weights=None, integ="SIMPSONS", log10F=False):
    f, w = _f64(f), _f64(weights)
    fn = lib().oracle_ref_calculate_snr
    fn.restype = C.c_double
    return fn(curve.encode(), detector.encode(), method.encode(), C.byref(src), _p(f), f.size, integ.encode(), _p(w), int(bool(log10F)))


def phenomd_intermediates(src):
    out = np.zeros(11)
    lib().oracle_ref_phenomd_intermediates(C.byref(src), _p(out))
    return dict(zip(["M", "eta", "chirpmass", "chi_pn", "A0", "fRD", "fdamp", "f1", "f3", "f1_phase", "f2_phase"], out))


def match(data1, data2, psd, f):
    """match() of the reference (src/waveform_util.cpp:41-89)."""
    f, psd = _f64(f), _f64(psd)
    a, b = np.asarray(data1), np.asarray(data2)
    fn = lib().oracle_ref_match
    fn.restype = C.c_double
    return fn(_p(_f64(a.real)), _p(_f64(a.imag)), _p(_f64(b.real)), _p(_f64(b.imag)), _p(psd), _p(f), f.size)


def mcmc_prep_params(method, mod, param, src):
    """MCMC_prep_params (src/mcmc_gw.cpp:2492-2568): returns (temp_params, the source with the flags and layout it set)."""
    param = _f64(param)
    temp = np.zeros_like(param)
    out = abi.Source()
    m = mod if mod is not None else abi.mod_defaults()
    lib().oracle_ref_mcmc_prep_params(method.encode(), C.byref(m), param.size, _p(param), C.byref(src), _p(temp), C.byref(out))
    return temp, out


def repack_mcmc_intrinsic(method, mod, params, gmst):
    """MCMC_prep_params (mcmc_intrinsic) + repack_parameters("MCMC_" + method) for the intrinsic sampling sets (src/fisher.cpp:2308-2376)."""
    params = _f64(params)
    W, P = params.shape
    out = (abi.Source * W)()
    m = mod if mod is not None else abi.mod_defaults()
    lib().oracle_ref_repack_mcmc_intrinsic(method.encode(), C.byref(m), P, W, _p(params), C.c_double(gmst), out)
    return out


def autocorrelation_lengths(positions, begin=0, target=0.01, nthreads=0):
    """calc_ac_vals' lags (auto_corr_from_data_batch, one cumulative segment) and the windowed estimator behind them."""
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    n_chains, steps, dim = pos.shape
    ac = np.zeros((n_chains, dim), dtype=np.int32)
    tau = np.zeros((n_chains, dim))
    lib().oracle_ref_autocorrelation_lengths(n_chains, dim, steps, _p(pos), int(begin), C.c_double(target), int(nthreads),
                                             ac.ctypes.data_as(C.POINTER(C.c_int)), _p(tau))
    return ac, tau


def pack_local_mod_structure(min_dim, max_dim, status, waveform_extended, full_mod):
    """pack_local_mod_structure (src/mcmc_gw.cpp:3401-3476): (counts[4], indices[4][MAX_MOD]) of the local structure."""
    status = np.ascontiguousarray(status, dtype=np.int32)
    counts = np.zeros(4, dtype=np.int32)
    idx = np.zeros((4, 8), dtype=np.int32)
    ip = C.POINTER(C.c_int)
    lib().oracle_ref_pack_local_mod_structure(int(min_dim), int(max_dim), status.ctypes.data_as(ip), waveform_extended.encode(),
                                              C.byref(full_mod), counts.ctypes.data_as(ip), idx.ctypes.data_as(ip))
    return counts, idx


def transform_orientation_coords(method, detector, src):
    """transform_orientation_coords (src/waveform_util.cpp:1535-1595): the (incl_angle, psi) it derives from (theta_l, phi_l)."""
    incl, psi = C.c_double(), C.c_double()
    lib().oracle_ref_transform_orientation_coords(method.encode(), detector.encode(), C.byref(src), C.byref(incl), C.byref(psi))
    return incl.value, psi.value
