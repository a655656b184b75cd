"""TEST INFRASTRUCTURE ONLY -- the reference's OWN sampler (src/mcmc_sampler_internals.cpp, src/mcmc_sampler.cpp) and its OWN
standard priors (src/standardPriorLibrary.cpp), compiled unmodified into oracle/_ref/libgwat_ref.so, driven from Python.

The reference draws its random numbers sequentially from one gsl_rng per chain; the CUDA sampler uses counter-based draws
(Philox keyed by (step, chain, purpose), gw_analysis_tools_b200/csrc/gwat_sampler_math.h).  To run the two side by side the
gsl_rng stand-in of the oracle build is *scripted*: this module computes, for every chain, the sequence of uniforms and unit
normals the reference's code will ask for -- in the order mcmc_step / gaussian_step / diff_ev_step / fisher_step /
PTMCMC_MH_step_incremental / single_chain_swap ask for them -- from the same counters the device uses, and queues them
(oracle/sampler_driver.cpp).  If the order assumed here were not the reference's, a queue would run dry or be left with
unconsumed draws (both are reported by `results()["diag"]`) and the trajectories would part from the device's at once.

Eigen (absent from this image) is a stand-in as well: a Jacobi solver, checked on its own; for step-by-step trajectory
comparisons the eigen-system each Fisher refresh obtains is scripted too (eigenvectors are defined up to sign), see
`RefSampler.script`.

Only tests/ may import this.
"""
import ctypes as C
import math

import numpy as np

from gw_analysis_tools_b200 import abi
from oracle import gwat_ref

DRAW_TYPE_ACCEPT, DRAW_PICK, DRAW_NORMAL, DRAW_DE_SCALE, DRAW_SWAP, DRAW_SWAP_GATE = range(6)
_dp = C.POINTER(C.c_double)
NCT = 12


def _p(a):
    return a.ctypes.data_as(_dp)


def log_prior_batch(method, params, prior, n_mod):
    """logPriorStandard_{D,P,D_NRT,P_NRT}[_mod]::eval of the reference on W sampling vectors."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    W, P = params.shape
    out = np.empty(W)
    gwat_ref.lib().oracle_ref_log_prior_batch(int("Pv2" in method), int("NRT" in method), int(n_mod), C.byref(prior), P, W, _p(params), _p(out))
    return out


def eigen_standin(A):
    """(eigenvalues ascending, eigenvectors as rows) from the oracle build's Eigen stand-in."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    vals, vecs = np.empty(n), np.empty((n, n))
    gwat_ref.lib().oracle_ref_eigen_standin(n, _p(A), _p(vals), _p(vecs))
    return vals, vecs


def step_boundaries(T, fisher_exist, primed):
    """assign_probabilities, non-RJ (src/mcmc_sampler_internals.cpp:1196-1270) -> cumulative boundaries; only used to know which
    draws the reference will consume next, the reference computes its own."""
    p = [0., 0., 0., 0.]
    if not fisher_exist:
        p[0] = 1.
    elif not primed:
        p[3] = .1 + .8 / T
        p[0] = 1 - (p[1] + p[2] + p[3])
    else:
        p[1] = .7 - .4 / T
        p[3] = .2 + .5 / T
        p[0] = 1 - (p[1] + p[2] + p[3] + 0.)
    b0 = p[0]
    b1 = p[1] + b0
    b2 = p[2] + b1
    b3 = p[3] + b2
    return b0, b1, b2, b3


class RefSampler:
    """The reference's `sampler` struct set up as PTMCMC_MH_internal does, stepped by PTMCMC_MH_step_incremental."""

    def __init__(self, wl, temps, initial, prior, seed, n_rounds, draw_uniform2, normal_from, swp_freq=5, swap_rate=None, history_length=1000,
                 history_update=10, fisher_exist=True, fisher_update_number=200, check_stepsize_freq=50, initial_fisher=None):
        """initial_fisher(chain) -> (vals, vecs): the eigen-system the chain's first matrix (computed by assign_initial_pos at the
        initial position, src/mcmc_sampler_internals.cpp:3182-3212) is scripted to have; needed when fisher_exist."""
        self.lib = gwat_ref.lib()
        self.lib.oracle_sampler_create.restype = C.c_void_p
        self.wl = wl
        self.temps = np.ascontiguousarray(temps, dtype=np.float64)
        init = np.ascontiguousarray(initial, dtype=np.float64)
        self.C, self.P = init.shape
        self.seed = int(seed)
        self.swp_freq = int(swp_freq)
        self.swap_rate = 1. / swp_freq if swap_rate is None else float(swap_rate)
        self.H, self.history_update = int(history_length), int(history_update)
        self.fisher_exist = bool(fisher_exist)
        self.fisher_update_number = int(fisher_update_number)
        self.n_rounds = int(n_rounds)
        self.N = self.n_rounds * self.swp_freq + 1   # positions stored per chain: the initial one + one per step
        self.u2, self.normal_from = draw_uniform2, normal_from
        base = 15 if "Pv2" in wl.method else 11
        if "NRT" in wl.method and "Pv2" not in wl.method:
            base += 1 if prior.tidal_love else 2
        n_mod = self.P - base
        f = np.ascontiguousarray(wl.f, dtype=np.float64)
        psd = np.ascontiguousarray(wl.psd, dtype=np.float64)
        dre = np.ascontiguousarray(wl.data.real, dtype=np.float64)
        dim = np.ascontiguousarray(wl.data.imag, dtype=np.float64)
        dets = (C.c_char_p * wl.D)(*[d.encode() for d in wl.detectors])
        ifv = ifw = None
        if self.fisher_exist:
            assert initial_fisher is not None, "a sampler with Fisher steps needs the chains' first eigen-systems"
            sys0 = [initial_fisher(c) for c in range(self.C)]
            ifv = np.ascontiguousarray([v for v, _ in sys0], dtype=np.float64)
            ifw = np.ascontiguousarray([w for _, w in sys0], dtype=np.float64)
        self._keep = (f, psd, dre, dim, dets, init, ifv, ifw)
        self.h = C.c_void_p(self.lib.oracle_sampler_create(
            wl.method.encode(), C.byref(wl.mod) if wl.mod is not None else None, int("Pv2" in wl.method), int("NRT" in wl.method), int(n_mod),
            self.C, self.P, self.N, _p(self.temps), _p(init), self.swp_freq, C.c_double(self.swap_rate), self.H, self.history_update,
            int(self.fisher_exist), self.fisher_update_number, int(check_stepsize_freq), C.byref(prior), C.c_double(wl.gmst),
            C.c_double(wl.T_segment), wl.D, dets, _p(f), f.size, _p(psd), _p(dre), _p(dim), _p(ifv) if ifv is not None else None,
            _p(ifw) if ifw is not None else None))

    def close(self):
        if self.h:
            self.lib.oracle_sampler_destroy(self.h)
            self.h = None

    def script(self, fisher_after_step=None):
        """Queue every draw of the whole run.  fisher_after_step(step, chain) -> (vals[P], vecs[P][P]): the eigen-system the
        chain's Fisher refresh at that step obtains (the device's, for step-by-step comparisons).  Returns, per chain, the list
        of (step, kind) with kind in 'gauss' | 'de' | 'fisher', and the list of (step, chain) at which refreshes happen."""
        C_, P, H = self.C, self.P, self.H
        # (assign_initial_pos has already asked for every chain's first matrix, at create: `initial_fisher`; counters run from 0)
        fisher_ct = [0] * C_
        kinds = [[] for _ in range(C_)]
        refreshes = []
        sweep = 0
        for rnd in range(self.n_rounds):
            for c in range(C_):
                u, n = [], []
                T = self.temps[c]
                for k in range(self.swp_freq):
                    s = rnd * self.swp_freq + k
                    primed = s > H
                    b = step_boundaries(T, self.fisher_exist, primed)
                    alpha, u_acc = self.u2(self.seed, s, c, DRAW_TYPE_ACCEPT)
                    u_pick, u_pick2 = self.u2(self.seed, s, c, DRAW_PICK)
                    z = self.normal_from(*self.u2(self.seed, s, c, DRAW_NORMAL))
                    u.append(alpha)
                    if alpha < b[0]:      # gaussian_step: dimension, then the jump
                        kinds[c].append((s, "gauss"))
                        u.append(u_pick)
                        n.append(z)
                    elif alpha < b[1]:    # diff_ev_step: i, j (redrawn until != i), scale gate, scale
                        kinds[c].append((s, "de"))
                        i = int(H * u_pick)
                        j = (i + 1 + int((H - 1) * u_pick2)) % H
                        beta, _ = self.u2(self.seed, s, c, DRAW_DE_SCALE)
                        u += [u_pick, (j + 0.5) / H, beta]
                        if beta < .9:
                            n.append(z)
                    else:                 # fisher_step: [refresh], eigen-direction, the jump
                        kinds[c].append((s, "fisher"))
                        if fisher_ct[c] == self.fisher_update_number:
                            refreshes.append((s, c))
                            if fisher_after_step is not None:
                                vals, vecs = fisher_after_step(s, c)
                                vals = np.ascontiguousarray(vals, dtype=np.float64)
                                vecs = np.ascontiguousarray(vecs, dtype=np.float64)
                                self.lib.oracle_sampler_push_fisher(self.h, c, _p(vals), _p(vecs))
                            fisher_ct[c] = 0
                        u.append(u_pick)
                        n.append(z)
                        fisher_ct[c] += 1
                    u.append(u_acc)       # mcmc_step: the Metropolis-Hastings draw
                self._push(c, u, n)
            # PTMCMC_MH_step_incremental: the sweep gate from chain 0's generator, then one draw per adjacent pair with different
            # temperatures from the lower chain's generator (single_chain_swap)
            gate, _ = self.u2(self.seed, sweep, 0, DRAW_SWAP_GATE)
            self._push(0, [gate], [])
            if gate < self.swap_rate:
                for i in range(C_ - 1):
                    if self.temps[i] != self.temps[i + 1]:
                        a, _ = self.u2(self.seed, sweep, i, DRAW_SWAP)
                        self._push(i, [a], [])
            sweep += 1
        return kinds, refreshes

    def _push(self, c, u, n):
        ua = np.ascontiguousarray(u, dtype=np.float64)
        na = np.ascontiguousarray(n, dtype=np.float64)
        self.lib.oracle_sampler_push(self.h, int(c), ua.size, _p(ua), na.size, _p(na))

    def run(self):
        rc = self.lib.oracle_sampler_run(self.h)
        assert rc == 0

    def results(self):
        C_, P, N = self.C, self.P, self.N
        out, llp = np.empty((C_, N, P)), np.empty((C_, N, 2))
        ct = np.zeros((C_, NCT), dtype=np.int64)
        widths, fvals, fvecs = np.empty((C_, P + 3)), np.empty((C_, P)), np.empty((C_, P, P))
        pos = np.zeros(C_, dtype=np.int32)
        diag = np.zeros(8, dtype=np.int64)
        self.lib.oracle_sampler_results(self.h, _p(out), _p(llp), ct.ctypes.data_as(C.POINTER(C.c_longlong)), _p(widths), _p(fvals), _p(fvecs),
                                        pos.ctypes.data_as(C.POINTER(C.c_int)), diag.ctypes.data_as(C.POINTER(C.c_longlong)))
        names = ["step_accept", "step_reject", "gauss_accept", "gauss_reject", "de_accept", "de_reject", "fisher_accept", "fisher_reject",
                 "swap_accept", "swap_reject", "fisher_updates", "fisher_nan"]
        return {"output": out, "ll": llp[:, :, 0], "lp": llp[:, :, 1], "counters": {n: ct[:, i] for i, n in enumerate(names)}, "widths": widths,
                "fvals": fvals, "fvecs": fvecs, "chain_pos": pos,
                "diag": {"rng_underflow": int(diag[0]), "uniforms_left": int(diag[1]), "normals_left": int(diag[2]),
                         "fisher_script_underflow": int(diag[3]), "ll_calls": int(diag[4]), "fisher_calls": int(diag[5])}}


def update_temperatures(chain_temps, A, t0, nu, t):
    """update_temperatures_full_ensemble of the reference (linear swapping, src/mcmc_sampler_internals.cpp:3371-3413)."""
    temps = np.array(chain_temps, dtype=np.float64)
    a = np.zeros(temps.size + 1, dtype=np.int32)
    a[:len(A)] = A
    gwat_ref.lib().oracle_ref_update_temperatures(int(temps.size), _p(temps), a.ctypes.data_as(C.POINTER(C.c_int)), int(t0), int(nu), int(t))
    return temps
