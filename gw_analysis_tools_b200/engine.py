"""ctypes binding of the product library ``libgwat_b200.so`` (the C ABI of ``include/gwat_b200.h``).

This is plumbing only: every numerical operation happens in the CUDA kernels behind the C ABI.  There is no CPU path;
``load_library`` raises if the extension was not built and ``Context`` raises if no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GWAT_B200_LIB") or os.path.join(_HERE, "libgwat_b200.so")  # env override: kernel experiments only

_dp = C.POINTER(C.c_double)
_lib = None

EXPORTS = [
    "gwat_b200_abi_version", "gwat_b200_cosmology_index", "gwat_b200_detector_site", "gwat_b200_transform_orientation_coords", "gwat_b200_source_init", "gwat_b200_mod_init", "gwat_b200_ctx_create",
    "gwat_b200_ctx_destroy", "gwat_b200_last_error", "gwat_b200_set_network", "gwat_b200_loglike_mcmc_batch",
    "gwat_b200_loglike_mcmc_batch_dev", "gwat_b200_loglike_batch", "gwat_b200_loglike_maximized_batch", "gwat_b200_loglike_maximized_mcmc_batch",
    "gwat_b200_fourier_waveform_batch", "gwat_b200_fourier_amplitude_phase_batch",
    "gwat_b200_coherent_response_batch", "gwat_b200_fourier_detector_response_batch",
    "gwat_b200_fisher_numerical_batch", "gwat_b200_repack_mcmc_batch", "gwat_b200_repack_mcmc_intrinsic_batch", "gwat_b200_antenna_batch",
    "gwat_b200_snr_batch", "gwat_b200_populate_noise", "gwat_b200_losc_prepare", "gwat_b200_gps_to_gmst_radian",
    "gwat_b200_queue_create", "gwat_b200_queue_destroy", "gwat_b200_queue_loglike", "gwat_b200_queue_stats",
    "gwat_b200_gauss_legendre_grid", "gwat_b200_log_likelihood_internal", "gwat_b200_match", "gwat_b200_method_info", "gwat_b200_measure_fp64_peak", "gwat_b200_set_kernel_timing", "gwat_b200_launch_count", "gwat_b200_last_kernel_ms", "gwat_b200_last_active_bins",
]


class GwatB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("gwat_b200 error %d: %s" % (code, message))
        self.code = code


def load_library():
    """Load the CUDA extension.  Fails loudly when it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("gwat_b200 CUDA extension missing: %s (run __graft_entry__.build()); there is no CPU fallback"
                              % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.gwat_b200_last_error.restype = C.c_char_p
        lib.gwat_b200_last_error.argtypes = [C.c_void_p]
        lib.gwat_b200_launch_count.restype = C.c_longlong
        lib.gwat_b200_launch_count.argtypes = [C.c_void_p]
        lib.gwat_b200_last_active_bins.restype = C.c_longlong
        lib.gwat_b200_last_active_bins.argtypes = [C.c_void_p]
        lib.gwat_b200_last_kernel_ms.restype = C.c_double
        lib.gwat_b200_last_kernel_ms.argtypes = [C.c_void_p]
        lib.gwat_b200_ctx_destroy.argtypes = [C.c_void_p]
        lib.gwat_b200_ctx_destroy.restype = None
        lib.gwat_b200_gps_to_gmst_radian.argtypes = [C.c_double]
        lib.gwat_b200_gps_to_gmst_radian.restype = C.c_double
        lib.gwat_b200_queue_destroy.argtypes = [C.c_void_p]
        lib.gwat_b200_queue_destroy.restype = None
        lib.gwat_b200_queue_loglike.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int)]
        lib.gwat_b200_queue_loglike.restype = C.c_double
        if lib.gwat_b200_abi_version() != abi.ABI_VERSION:
            raise ImportError("gwat_b200 ABI mismatch")
        _lib = lib
    return _lib


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _src_array(sources):
    if isinstance(sources, abi.Source):
        sources = [sources]
    if isinstance(sources, C.Array):
        return sources, len(sources)
    return (abi.Source * len(sources))(*sources), len(sources)


def gauss_legendre_grid(f_lower, f_upper, n, log10F=True):
    """(frequencies, weights) of the reference's GAUSSLEG grids (host code of the library)."""
    f, w = np.empty(n), np.empty(n)
    rc = load_library().gwat_b200_gauss_legendre_grid(C.c_double(f_lower), C.c_double(f_upper), int(n), int(bool(log10F)), _p(f), _p(w))
    if rc != 0:
        raise GwatB200Error(rc, "gauss_legendre_grid: bad arguments")
    return f, w


def gps_to_gmst_radian(gps_time):
    """Greenwich mean sidereal time (rad) of a GPS time, by the reference's formula."""
    return load_library().gwat_b200_gps_to_gmst_radian(float(gps_time))


def populate_noise(frequencies, curve, noise_data_dir=None):
    """sqrt(S_n(f)) of a named GWAT noise curve (host code of the library; tabulated curves are read from ``noise_data_dir``)."""
    f = _f64(frequencies)
    out = np.empty(f.size)
    rc = load_library().gwat_b200_populate_noise(_p(f), curve.encode(), noise_data_dir.encode() if noise_data_dir else None, int(f.size),
                                                 _p(out))
    if rc != 0:
        raise GwatB200Error(rc, "populate_noise(%r): unknown curve, unreadable file or frequency outside the table" % curve)
    return out


class Context:
    """One GPU's worth of state: the detector network, the frequency grid and scratch buffers."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.gwat_b200_ctx_create(C.byref(self._h), int(device))
        if rc != 0:
            raise GwatB200Error(rc, self._lib.gwat_b200_last_error(None).decode())
        self.device = int(device)
        self.D = 0
        self.L = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.gwat_b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise GwatB200Error(rc, self._lib.gwat_b200_last_error(self._h).decode())

    # ---- network ---------------------------------------------------------------------------------------------------
    def set_network(self, detectors, frequencies, psd, data=None, weights=None, integration_method="SIMPSONS",
                    log10F=False):
        f = _f64(frequencies)
        psd = _f64(psd)
        D, L = len(detectors), f.size
        assert psd.shape == (D, L), "psd must have shape [D][L]"
        dre = dim = None
        if data is not None:
            data = np.asarray(data)
            assert data.shape == (D, L)
            dre, dim = _f64(data.real), _f64(data.imag)
        w = _f64(weights)
        names = (C.c_char_p * D)(*[d.encode() for d in detectors])
        self._check(self._lib.gwat_b200_set_network(self._h, D, names, L, _p(f), _p(psd), _p(dre), _p(dim), _p(w),
                                                    integration_method.encode(), int(bool(log10F))))
        self.D, self.L = D, L

    # ---- likelihood ------------------------------------------------------------------------------------------------
    def loglike_mcmc_batch(self, method, params, gmst, T_segment, mod=None):
        params = _f64(params)
        W, P = params.shape
        out = np.empty(W)
        self._check(self._lib.gwat_b200_loglike_mcmc_batch(self._h, method.encode(), C.byref(mod) if mod is not None else None,
                                                           P, W, _p(params), C.c_double(gmst), C.c_double(T_segment), _p(out)))
        return out

    def loglike_maximized_batch(self, method, sources):
        """tc/phic-maximised log-likelihoods (the reference's intrinsic samplers), one per source."""
        arr, W = _src_array(sources)
        out = np.empty(W)
        self._check(self._lib.gwat_b200_loglike_maximized_batch(self._h, method.encode(), W, arr, _p(out)))
        return out

    def loglike_mcmc_batch_dev(self, method, d_params_ptr, W, P, gmst, T_segment, d_logL_ptr, mod=None, stream=None):
        """Device-pointer variant (integers from ``tensor.data_ptr()``); asynchronous on ``stream`` (a raw cudaStream_t)."""
        self._check(self._lib.gwat_b200_loglike_mcmc_batch_dev(
            self._h, method.encode(), C.byref(mod) if mod is not None else None, int(P), int(W), C.c_void_p(d_params_ptr),
            C.c_double(gmst), C.c_double(T_segment), C.c_void_p(d_logL_ptr), C.c_void_p(stream or 0)))

    def losc_prepare(self, data_files, psd_file, trigger_time, post_merger_duration):
        """(frequencies[L], psd[D][L], data[D][L]) from LOSC strain text files and a LOSC PSD file (the reference's allocate_LOSC_data)."""
        D = len(data_files)
        names = (C.c_char_p * D)(*[str(p).encode() for p in data_files])
        n = C.c_int(0)
        rc = self._lib.gwat_b200_losc_prepare(self._h, D, names, str(psd_file).encode(), C.c_double(trigger_time),
                                              C.c_double(post_merger_duration), 0, C.byref(n), None, None, None, None)
        if n.value <= 0:
            self._check(rc)
        L = n.value
        f, psd, dre, dim = np.empty(L), np.empty((D, L)), np.empty((D, L)), np.empty((D, L))
        self._check(self._lib.gwat_b200_losc_prepare(self._h, D, names, str(psd_file).encode(), C.c_double(trigger_time),
                                                     C.c_double(post_merger_duration), L, C.byref(n), _p(f), _p(psd), _p(dre), _p(dim)))
        return f, psd, dre + 1j * dim

    def snr_batch(self, method, sources):
        """Network matched-filter SNR of each source (one-detector network: the reference's calculate_snr)."""
        arr, W = _src_array(sources)
        out = np.empty(W)
        self._check(self._lib.gwat_b200_snr_batch(self._h, method.encode(), W, arr, _p(out)))
        return out

    def loglike_batch(self, method, sources):
        arr, W = _src_array(sources)
        out = np.empty(W)
        self._check(self._lib.gwat_b200_loglike_batch(self._h, method.encode(), W, arr, _p(out)))
        return out

    # ---- waveforms / responses -----------------------------------------------------------------------------------------
    def fourier_waveform_batch(self, method, sources):
        arr, W = _src_array(sources)
        o = [np.empty((W, self.L)) for _ in range(4)]
        self._check(self._lib.gwat_b200_fourier_waveform_batch(self._h, method.encode(), W, arr, *[_p(x) for x in o]))
        return o[0] + 1j * o[1], o[2] + 1j * o[3]

    def fourier_amplitude_phase_batch(self, method, sources):
        """(amplitude[W][L], phase[W][L]) of the IMRPhenomD-family carrier: the reference's fourier_amplitude / fourier_phase."""
        arr, W = _src_array(sources)
        a, p = np.empty((W, self.L)), np.empty((W, self.L))
        self._check(self._lib.gwat_b200_fourier_amplitude_phase_batch(self._h, method.encode(), W, arr, _p(a), _p(p)))
        return a, p

    def coherent_response_batch(self, method, sources):
        arr, W = _src_array(sources)
        re, im = np.empty((W, self.D, self.L)), np.empty((W, self.D, self.L))
        self._check(self._lib.gwat_b200_coherent_response_batch(self._h, method.encode(), W, arr, _p(re), _p(im)))
        return re + 1j * im

    def fourier_detector_response_batch(self, method, detector, sources):
        arr, W = _src_array(sources)
        re, im = np.empty((W, self.L)), np.empty((W, self.L))
        self._check(self._lib.gwat_b200_fourier_detector_response_batch(self._h, method.encode(), detector.encode(), W, arr,
                                                                        _p(re), _p(im)))
        return re + 1j * im

    def fisher_numerical_batch(self, method, sources, dimension, order=4, detector_index=-1, reference_index=0):
        arr, S = _src_array(sources)
        out = np.empty((S, dimension, dimension))
        self._check(self._lib.gwat_b200_fisher_numerical_batch(self._h, method.encode(), int(detector_index),
                                                               int(reference_index), int(dimension), int(order), S, arr,
                                                               _p(out)))
        return out

    def log_likelihood_internal(self, data, psd, frequencies, response, weights=None, integration_method="SIMPSONS", log10F=False):
        """The reference's Log_Likelihood_internal for one detector and a response held in host memory."""
        f, psd, w = _f64(frequencies), _f64(psd), _f64(weights)
        dre, dim = _f64(np.asarray(data).real), _f64(np.asarray(data).imag)
        rre, rim = _f64(np.asarray(response).real), _f64(np.asarray(response).imag)
        out = C.c_double()
        self._check(self._lib.gwat_b200_log_likelihood_internal(self._h, int(f.size), _p(f), _p(psd), _p(dre), _p(dim), _p(w),
                                                                integration_method.encode(), int(bool(log10F)), _p(rre), _p(rim), C.byref(out)))
        return out.value

    def match(self, data1, data2, psd, frequencies):
        """The reference's match(): overlap of two frequency series maximised over a relative time shift."""
        f, psd = _f64(frequencies), _f64(psd)
        a, b = np.asarray(data1), np.asarray(data2)
        out = C.c_double()
        self._check(self._lib.gwat_b200_match(self._h, int(f.size), _p(f), _p(psd), _p(_f64(a.real)), _p(_f64(a.imag)), _p(_f64(b.real)),
                                              _p(_f64(b.imag)), C.byref(out)))
        return out.value

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def repack_mcmc_batch(self, method, params, gmst, mod=None):
        params = _f64(params)
        W, P = params.shape
        out = (abi.Source * W)()
        self._check(self._lib.gwat_b200_repack_mcmc_batch(self._h, method.encode(), C.byref(mod) if mod is not None else None,
                                                          P, W, _p(params), C.c_double(gmst), out))
        return out

    def repack_mcmc_intrinsic_batch(self, method, params, gmst, mod=None):
        """Sampling vectors of the intrinsic sets (ln Mc, eta, spins [, tidal] [, modifications]) -> physical records."""
        params = _f64(params)
        W, P = params.shape
        out = (abi.Source * W)()
        self._check(self._lib.gwat_b200_repack_mcmc_intrinsic_batch(self._h, method.encode(), C.byref(mod) if mod is not None else None,
                                                                    P, W, _p(params), C.c_double(gmst), out))
        return out

    def loglike_maximized_mcmc_batch(self, method, params, gmst, mod=None):
        """The intrinsic branch of MCMC_likelihood_wrapper for W sampling vectors: tc/phic-maximised log-likelihoods."""
        params = _f64(params)
        W, P = params.shape
        out = np.empty(W)
        self._check(self._lib.gwat_b200_loglike_maximized_mcmc_batch(self._h, method.encode(), C.byref(mod) if mod is not None else None,
                                                                     P, W, _p(params), C.c_double(gmst), _p(out)))
        return out

    def antenna_batch(self, RA, DEC, psi, gmst):
        RA, DEC, psi = _f64(RA), _f64(DEC), _f64(psi)
        W = RA.size
        fp, fc, dt = np.empty((W, self.D)), np.empty((W, self.D)), np.empty((W, self.D))
        self._check(self._lib.gwat_b200_antenna_batch(self._h, W, _p(RA), _p(DEC), _p(psi), C.c_double(gmst), _p(fp), _p(fc),
                                                      _p(dt)))
        return fp, fc, dt

    # ---- introspection ---------------------------------------------------------------------------------------------
    def measure_fp64_peak(self):
        """Measured DFMA issue peak of this GPU in TFLOP/s (the FP64 roofline denominator)."""
        out = C.c_double()
        self._check(self._lib.gwat_b200_measure_fp64_peak(self._h, C.byref(out)))
        return out.value

    @property
    def launch_count(self):
        return int(self._lib.gwat_b200_launch_count(self._h))

    def set_kernel_timing(self, on):
        """Record CUDA events around the hot likelihood kernel of every call (off by default: they cost ~6 us per call)."""
        self._check(self._lib.gwat_b200_set_kernel_timing(self._h, int(bool(on))))

    @property
    def last_kernel_ms(self):
        return float(self._lib.gwat_b200_last_kernel_ms(self._h))

    @property
    def last_active_bins(self):
        return int(self._lib.gwat_b200_last_active_bins(self._h))


class LikelihoodQueue:
    """One-chain-per-call likelihood with the reference callback's shape; concurrent calls are merged into batched launches
    (``gwat_b200_queue_*``).  ``loglike`` blocks and may be called from many threads (ctypes releases the GIL)."""

    def __init__(self, ctx, method, dimension, gmst, T_segment, mod=None, max_batch=4096, expected_callers=16, max_wait_us=200.0):
        self._lib = ctx._lib
        self._ctx = ctx  # keeps the context alive
        self._h = C.c_void_p()
        rc = self._lib.gwat_b200_queue_create(C.byref(self._h), ctx._h, method.encode(), C.byref(mod) if mod is not None else None,
                                              int(dimension), C.c_double(gmst), C.c_double(T_segment), int(max_batch),
                                              int(expected_callers), C.c_double(max_wait_us))
        if rc != 0:
            raise GwatB200Error(rc, "queue_create: bad arguments")
        self.dimension = int(dimension)

    def loglike(self, param):
        param = _f64(param)
        assert param.shape == (self.dimension,)
        status = C.c_int()
        v = self._lib.gwat_b200_queue_loglike(self._h, _p(param), C.byref(status))
        if status.value != 0:
            raise GwatB200Error(status.value, self._lib.gwat_b200_last_error(self._ctx._h).decode())
        return v

    def stats(self):
        calls, batches, largest = C.c_longlong(), C.c_longlong(), C.c_int()
        self._lib.gwat_b200_queue_stats(self._h, C.byref(calls), C.byref(batches), C.byref(largest))
        return calls.value, batches.value, largest.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.gwat_b200_queue_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
