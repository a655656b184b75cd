"""ctypes binding of the device-resident parallel-tempering sampler (``include/gwat_b200_sampler.h``).

Plumbing only, like ``engine.py``: proposals, priors, likelihoods, acceptance and swaps all run in CUDA kernels.
"""
import ctypes as C

import numpy as np

from . import abi
from .engine import Context, GwatB200Error, _f64, _p  # noqa: F401

_d2 = C.c_double * 2

COUNTERS = ["step_accept", "step_reject", "gauss_accept", "gauss_reject", "de_accept", "de_reject", "fisher_accept",
            "fisher_reject", "swap_accept", "swap_reject", "fisher_updates", "fisher_nan"]

EXPORTS = [
    "gwat_b200_prior_init", "gwat_b200_sampler_options_init", "gwat_b200_sampler_create", "gwat_b200_sampler_destroy",
    "gwat_b200_sampler_run", "gwat_b200_sampler_state", "gwat_b200_sampler_counters", "gwat_b200_sampler_cold",
    "gwat_b200_sampler_fisher_state", "gwat_b200_sampler_set_state", "gwat_b200_swap_sweep_host", "gwat_b200_swap_sweep_device", "gwat_b200_sampler_uniform",
    "gwat_b200_sampler_last_ms", "gwat_b200_sampler_last_launches", "gwat_b200_log_prior_batch", "gwat_b200_mcmc_fisher_batch", "gwat_b200_mcmc_fisher_intrinsic_batch",
    "gwat_b200_nccl_unique_id", "gwat_b200_sampler_attach_ranks", "gwat_b200_sampler_last_swap_ms", "gwat_b200_sampler_last_sweeps",
    "gwat_b200_update_temperatures", "gwat_b200_sampler_set_temperatures", "gwat_b200_sampler_temperatures",
    "gwat_b200_sampler_last_swap_accepts", "gwat_b200_sampler_dynamic_temperatures",
    # chain output (bound in chain_io.py)
    "gwat_b200_dump_create", "gwat_b200_dump_write", "gwat_b200_dump_close", "gwat_b200_write_data_dump", "gwat_b200_write_flat_thin_output",
    "gwat_b200_autocorrelation_lengths",
]
NCCL_UNIQUE_ID_BYTES = 128


def nccl_unique_id():
    """128 opaque bytes identifying a new group of ranks (rank 0 calls this and hands them to the others)."""
    from .engine import load_library
    buf = (C.c_ubyte * NCCL_UNIQUE_ID_BYTES)()
    lib = load_library()
    rc = lib.gwat_b200_nccl_unique_id(buf)
    if rc != 0:
        raise GwatB200Error(rc, lib.gwat_b200_last_error(None).decode())
    return bytes(buf)


class Prior(C.Structure):
    """``gwat_b200_prior`` = the fields of ``priorData`` (include/gwat/standardPriorLibrary.h:8-35) the standard priors read."""

    _fields_ = [(n, _d2) for n in ("mass1_prior", "mass2_prior", "spin1_prior", "spin2_prior", "a1_prior", "a2_prior", "ctheta1_prior",
                                   "ctheta2_prior", "phi1_prior", "phi2_prior", "tidal1_prior", "tidal2_prior", "tidal_s_prior",
                                   "RA_bounds", "sinDEC_bounds", "DL_prior")] + [
        ("T_merger", C.c_double), ("mod_priors", _d2 * abi.MAX_MOD), ("tidal_love", C.c_int), ("reserved_", C.c_int)]

    def as_dict(self):
        d = {n: list(getattr(self, n)) for n, t in self._fields_ if t is _d2}
        d["T_merger"] = self.T_merger
        d["tidal_love"] = self.tidal_love
        d["mod_priors"] = [list(r) for r in self.mod_priors]
        return d


class Options(C.Structure):
    _fields_ = [("chain_N", C.c_int), ("dimension", C.c_int), ("swp_freq", C.c_int), ("swap_rate", C.c_double),
                ("history_length", C.c_int), ("history_update", C.c_int), ("fisher_exist", C.c_int),
                ("fisher_update_number", C.c_int), ("fisher_deriv_order", C.c_int), ("check_stepsize_freq", C.c_int),
                ("seed", C.c_ulonglong), ("lanes", C.c_int), ("record_cold", C.c_int), ("fisher_deferred", C.c_int), ("chain_index_offset", C.c_int), ("reserved_", C.c_int)]


def prior_defaults(**kw):
    from .engine import load_library
    p = Prior()
    load_library().gwat_b200_prior_init(C.byref(p))
    for k, v in kw.items():
        cur = getattr(p, k)
        if k == "mod_priors":
            for i, (lo, hi) in enumerate(v):
                cur[i][0], cur[i][1] = lo, hi
        elif hasattr(cur, "__len__"):
            cur[0], cur[1] = v
        else:
            setattr(p, k, v)
    return p


def options_defaults(**kw):
    from .engine import load_library
    o = Options()
    load_library().gwat_b200_sampler_options_init(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    if "swp_freq" in kw and "swap_rate" not in kw:
        o.swap_rate = 1.0 / o.swp_freq  # src/mcmc_sampler.cpp:4316
    return o


def log_prior_batch(ctx, method, params, prior, mod=None):
    params = _f64(params)
    W, P = params.shape
    out = np.empty(W)
    ctx._check(ctx._lib.gwat_b200_log_prior_batch(ctx._h, method.encode(), C.byref(mod) if mod is not None else None, P, W,
                                                  C.byref(prior), _p(params), _p(out)))
    return out


def mcmc_fisher_batch(ctx, method, params, gmst, order=4, mod=None):
    """Fisher matrices as the reference's ``MCMC_fisher_wrapper`` returns them, and their eigen-systems."""
    params = _f64(params)
    W, P = params.shape
    F, vals, vecs = np.empty((W, P, P)), np.empty((W, P)), np.empty((W, P, P))
    ctx._check(ctx._lib.gwat_b200_mcmc_fisher_batch(ctx._h, method.encode(), C.byref(mod) if mod is not None else None, P, int(order), W,
                                                    _p(params), C.c_double(gmst), _p(F), _p(vals), _p(vecs)))
    return F, vals, vecs


def mcmc_fisher_intrinsic_batch(ctx, method, params, gmst, order=4, mod=None):
    """``MCMC_fisher_wrapper`` of an intrinsic run (ln Mc, eta, chi1, chi2 [, modifications]): sum over detectors + transformations."""
    params = _f64(params)
    W, P = params.shape
    F = np.empty((W, P, P))
    ctx._check(ctx._lib.gwat_b200_mcmc_fisher_intrinsic_batch(ctx._h, method.encode(), C.byref(mod) if mod is not None else None, P, int(order), W,
                                                              _p(params), C.c_double(gmst), _p(F)))
    return F


def autocorrelation_lengths(ctx, positions, begin=0):
    """(ac_values, tau) of ``positions[n_chains][steps][dimension]`` from step ``begin`` on: the lags ``calc_ac_vals`` feeds the thinning
    with (emcee's windowed estimator, truncated), and the estimator itself."""
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    n_chains, steps, dim = pos.shape
    ac = np.zeros((n_chains, dim), dtype=np.int32)
    tau = np.zeros((n_chains, dim))
    ctx._check(ctx._lib.gwat_b200_autocorrelation_lengths(ctx._h, n_chains, dim, C.c_longlong(steps), _p(pos), C.c_longlong(begin),
                                                          ac.ctypes.data_as(C.POINTER(C.c_int)), _p(tau)))
    return ac, tau


class Sampler:
    """All chains of a PTMCMC run, resident on one GPU.  ``ctx`` must already hold the network and the data."""

    def __init__(self, ctx, method, temps, initial_positions, prior, gmst, T_segment, mod=None, **options):
        self._ctx = ctx
        self._lib = ctx._lib
        init = _f64(initial_positions)
        temps = _f64(temps)
        self.C, self.P = init.shape
        assert temps.shape == (self.C,)
        self.options = options_defaults(chain_N=self.C, dimension=self.P, **options)
        self._lib.gwat_b200_sampler_last_ms.restype = C.c_double
        self._lib.gwat_b200_sampler_last_ms.argtypes = [C.c_void_p]
        self._lib.gwat_b200_sampler_last_launches.restype = C.c_longlong
        self._lib.gwat_b200_sampler_last_launches.argtypes = [C.c_void_p]
        self._lib.gwat_b200_sampler_destroy.argtypes = [C.c_void_p]
        self._lib.gwat_b200_sampler_destroy.restype = None
        self._h = C.c_void_p()
        ctx._check(self._lib.gwat_b200_sampler_create(ctx._h, method.encode(), C.byref(mod) if mod is not None else None,
                                                      C.byref(self.options), C.byref(prior), _p(temps), _p(init), C.c_double(gmst),
                                                      C.c_double(T_segment), C.byref(self._h)))
        self.steps = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.gwat_b200_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n_steps):
        self._ctx._check(self._lib.gwat_b200_sampler_run(self._h, int(n_steps)))
        self.steps += int(n_steps)

    def state(self):
        pos, ll, lp = np.empty((self.C, self.P)), np.empty(self.C), np.empty(self.C)
        self._ctx._check(self._lib.gwat_b200_sampler_state(self._h, _p(pos), _p(ll), _p(lp)))
        return pos, ll, lp

    def set_state(self, positions=None, logL=None, logP=None):
        a = [None if x is None else _f64(x) for x in (positions, logL, logP)]
        self._ctx._check(self._lib.gwat_b200_sampler_set_state(self._h, *[_p(x) for x in a]))

    def counters(self):
        ct = np.empty((self.C, len(COUNTERS)), dtype=np.int64)
        w = np.empty((self.C, self.P + 3))
        self._ctx._check(self._lib.gwat_b200_sampler_counters(self._h, ct.ctypes.data_as(C.POINTER(C.c_longlong)), _p(w)))
        return {n: ct[:, i] for i, n in enumerate(COUNTERS)}, w

    def fisher_state(self):
        vals, vecs = np.empty((self.C, self.P)), np.empty((self.C, self.P, self.P))
        self._ctx._check(self._lib.gwat_b200_sampler_fisher_state(self._h, _p(vals), _p(vecs)))
        return vals, vecs

    def cold_chains(self, first_step=0, n=None):
        nc = C.c_int()
        self._ctx._check(self._lib.gwat_b200_sampler_cold(self._h, C.c_longlong(0), 0, None, C.byref(nc)))
        n = self.steps - first_step if n is None else n
        out = np.empty((n, nc.value, self.P))
        self._ctx._check(self._lib.gwat_b200_sampler_cold(self._h, C.c_longlong(first_step), int(n), _p(out), C.byref(nc)))
        return out

    # ---- dynamic temperature allocation (arXiv:1501.05823; src/mcmc_sampler.cpp:453-545) ----
    def temperatures(self):
        out = np.empty(self.C)
        self._ctx._check(self._lib.gwat_b200_sampler_temperatures(self._h, _p(out)))
        return out

    def set_temperatures(self, temps):
        t = np.ascontiguousarray(temps, dtype=np.float64)
        self._ctx._check(self._lib.gwat_b200_sampler_set_temperatures(self._h, _p(t)))

    def last_swap_accepts(self):
        out = np.zeros(max(self.C - 1, 1), dtype=np.int32)
        self._ctx._check(self._lib.gwat_b200_sampler_last_swap_accepts(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out[:self.C - 1]

    def dynamic_temperatures(self, n_steps, nu=10, t0=1000):
        """Tune the ladder for n_steps steps (blocks of swp_freq steps, a sweep and a temperature update after each); returns the sweeps done."""
        n = C.c_longlong()
        self._ctx._check(self._lib.gwat_b200_sampler_dynamic_temperatures(self._h, int(n_steps), int(nu), int(t0), C.byref(n)))
        self.steps += n.value * self.options.swp_freq
        return n.value

    def attach_ranks(self, unique_id, rank, n_ranks):
        """Join a ladder sharded over ``n_ranks`` processes (one per GPU): this sampler must have been created with its own
        equal share of the chains and ``chain_index_offset = rank * chain_N``.  The swap sweeps then exchange over NCCL."""
        buf = (C.c_ubyte * NCCL_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._ctx._check(self._lib.gwat_b200_sampler_attach_ranks(self._h, buf, int(rank), int(n_ranks)))

    @property
    def last_swap_ms(self):
        self._lib.gwat_b200_sampler_last_swap_ms.restype = C.c_double
        self._lib.gwat_b200_sampler_last_swap_ms.argtypes = [C.c_void_p]
        return self._lib.gwat_b200_sampler_last_swap_ms(self._h)

    @property
    def last_sweeps(self):
        self._lib.gwat_b200_sampler_last_sweeps.restype = C.c_longlong
        self._lib.gwat_b200_sampler_last_sweeps.argtypes = [C.c_void_p]
        return self._lib.gwat_b200_sampler_last_sweeps(self._h)

    @property
    def last_ms(self):
        return self._lib.gwat_b200_sampler_last_ms(self._h)

    @property
    def last_launches(self):
        return self._lib.gwat_b200_sampler_last_launches(self._h)


def update_temperatures(chain_temps, A, t0, nu, t):
    """update_temperatures_full_ensemble with linear swapping (src/mcmc_sampler_internals.cpp:3371-3413): the new ladder."""
    from .engine import load_library
    temps = np.array(chain_temps, dtype=np.float64)
    a = np.zeros(temps.size + 1)
    a[:len(A)] = A
    rc = load_library().gwat_b200_update_temperatures(int(temps.size), _p(temps), _p(a), int(t0), int(nu), int(t))
    if rc != 0:
        raise ValueError("gwat_b200_update_temperatures: %d" % rc)
    return temps


def swap_sweep_host(logL, temps, seed, sweep):
    """The reference's swap sweep over a whole ladder (host code of the library, the single-GPU sweep's draws and arithmetic):
    returns (src, accepted) with new_state[i] = old_state[src[i]]."""
    from .engine import load_library
    logL, temps = _f64(logL), _f64(temps)
    n = logL.size
    src = np.empty(n, dtype=np.int32)
    acc = np.zeros(max(n - 1, 1), dtype=np.int32)
    rc = load_library().gwat_b200_swap_sweep_host(n, _p(logL), _p(temps), C.c_ulonglong(seed), C.c_longlong(sweep),
                                                  src.ctypes.data_as(C.POINTER(C.c_int)), acc.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != 0:
        raise GwatB200Error(rc, "swap_sweep_host: bad arguments")
    return src, acc[:n - 1]


def swap_sweep_device(ctx, logL, temps, seed, sweep, mode=0):
    """The sweep as the device sampler computes it (thresholds in parallel, run starts by pointer doubling; mode 1: the sequential
    fallback) on the given ladder: (src, accepted), identical to swap_sweep_host's."""
    from .engine import load_library
    logL, temps = _f64(logL), _f64(temps)
    n = logL.size
    src = np.empty(n, dtype=np.int32)
    acc = np.zeros(max(n - 1, 1), dtype=np.int32)
    lib = load_library()
    rc = lib.gwat_b200_swap_sweep_device(ctx._h, n, _p(logL), _p(temps), C.c_ulonglong(seed), C.c_longlong(sweep), int(mode),
                                         src.ctypes.data_as(C.POINTER(C.c_int)), acc.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != 0:
        raise GwatB200Error(rc, lib.gwat_b200_last_error(ctx._h).decode())
    return src, acc[:n - 1]


def draw_uniform2(seed, step, chain, purpose):
    from .engine import load_library
    out = (C.c_double * 2)()
    lib = load_library()
    lib.gwat_b200_sampler_uniform.restype = None
    lib.gwat_b200_sampler_uniform(C.c_ulonglong(seed), C.c_ulonglong(step), C.c_uint(chain), C.c_uint(purpose), out)
    return out[0], out[1]


def prior_for(wl):
    """Standard-prior bounds that contain a synthetic workload (``workloads.make``): masses 1-100 Msun (0.5-3 for the
    neutron-star config), D_L 1 Mpc - 10 Gpc, full spin and angle ranges, tc within 0.1 s of the injected value."""
    import math
    ns = "NRT" in wl.method
    return prior_defaults(
        mass1_prior=(0.5, 3.0) if ns else (1.0, 100.0), mass2_prior=(0.5, 3.0) if ns else (1.0, 100.0),
        spin1_prior=(-0.05, 0.05) if ns else (-0.95, 0.95), spin2_prior=(-0.05, 0.05) if ns else (-0.95, 0.95),
        a1_prior=(0.0, 0.95), a2_prior=(0.0, 0.95), ctheta1_prior=(-1.0, 1.0), ctheta2_prior=(-1.0, 1.0),
        phi1_prior=(0.0, 2 * math.pi), phi2_prior=(0.0, 2 * math.pi), tidal1_prior=(1.0, 5000.0), tidal2_prior=(1.0, 5000.0),
        tidal_s_prior=(1.0, 5000.0), RA_bounds=(0.0, 2 * math.pi), sinDEC_bounds=(-1.0, 1.0), DL_prior=(1.0, 10000.0),
        T_merger=float(wl.inj[5]), tidal_love=1, mod_priors=[(0.0, 50.0)] * abi.MAX_MOD)
