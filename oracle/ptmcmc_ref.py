"""TEST INFRASTRUCTURE -- CPU restatement of the reference's parallel-tempering Metropolis-Hastings step.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this.  It restates, function by function,

    mcmc_step, gaussian_step, diff_ev_step, fisher_step, assign_probabilities (non-RJ), update_step_widths,
    update_history, chain_swap / single_chain_swap        /root/reference/src/mcmc_sampler_internals.cpp
    PTMCMC_MH_step_incremental (non-pool loop)            /root/reference/src/mcmc_sampler.cpp:4571-4660
    MCMC_fisher_transformations                           /root/reference/src/mcmc_gw.cpp:2136-2189
    logPriorStandard_{D,P,D_NRT,P_NRT}[_mod]::eval        /root/reference/src/standardPriorLibrary.cpp:321-526

in plain Python/numpy loops, one chain at a time like the reference, with the log-likelihood and Fisher matrix supplied by
the caller (the tests pass the compiled reference code from oracle/_ref).

PARITY UNPINNED for this file: the reference's sampler cannot be compiled here (it needs Eigen, GSL's generators and
BayesShip) and its tests store no trajectories.  What IS pinned: every likelihood/Fisher value the sampler consumes comes
from the compiled reference, and the rules below are short enough to be checked against the cited lines by eye.

Random numbers: the reference draws from one gsl_rng per chain (mt19937 seeded with chain+1); a batched sampler cannot
consume a sequential stream, so both this file and the CUDA path use Philox4x32-10 with counter (step, chain, purpose).
Which draw feeds which decision is fixed by `Draws` below; the distribution of every decision is the reference's.
"""
import math

import numpy as np

M32 = 0xFFFFFFFF
DRAW_TYPE_ACCEPT, DRAW_PICK, DRAW_NORMAL, DRAW_DE_SCALE, DRAW_SWAP, DRAW_SWAP_GATE = range(6)
STEP_GAUSS, STEP_DE, STEP_MMALA, STEP_FISHER = range(4)


def philox4x32_10(counter, key):
    """Salmon, Moraes, Dror, Shaw (SC'11): Philox-4x32 with 10 rounds.  counter: 4 words, key: 2 words."""
    c0, c1, c2, c3 = [int(x) & M32 for x in counter]
    k0, k1 = [int(x) & M32 for x in key]
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def uniform2(seed, step, chain, purpose):
    r = philox4x32_10((step & M32, (step >> 32) & M32, chain, purpose), (seed & M32, (seed >> 32) & M32))
    u0 = ((r[0] >> 5) * 67108864.0 + (r[1] >> 6)) / 9007199254740992.0
    u1 = ((r[2] >> 5) * 67108864.0 + (r[3] >> 6)) / 9007199254740992.0
    return u0, u1


def normal_from(u0, u1):
    return math.sqrt(-2.0 * math.log(1.0 - u0)) * math.cos(2 * math.pi * u1)


# ---- priors: src/standardPriorLibrary.cpp ---------------------------------------------------------------------------------

def _csqrt(x):  # C semantics: NaN instead of an exception
    return math.sqrt(x) if x >= 0 else math.nan


def _cpow(x, y):
    return x ** y if x > 0 else (0.0 if x == 0 else math.nan)


def calculate_mass1(chirpmass, eta):  # src/util.cpp:1516-1520
    etapow = _cpow(eta, 3. / 5)
    if etapow == 0:
        return math.nan
    return 1. / 2 * (chirpmass / etapow + _csqrt(1. - 4 * eta) * chirpmass / etapow)


def calculate_mass2(chirpmass, eta):  # src/util.cpp:1531-1535
    etapow = _cpow(eta, 3. / 5)
    if etapow == 0:
        return math.nan
    return 1. / 2 * (chirpmass / etapow - _csqrt(1. - 4 * eta) * chirpmass / etapow)


def chirpmass_eta_jac(chirpmass, eta):  # :10-19
    epsilon = 1e-12
    delta = math.sqrt(1. - 4. * eta)
    if eta > .25 - epsilon:
        delta = math.sqrt(1. - 4. * (eta - epsilon))
    return chirpmass * chirpmass / (delta * eta ** 1.2)


def aligned_spin_prior(chi):  # :25-28
    return 0.0039132 * math.exp(-3.95381 * abs(chi))


def tidal_love_boundary_violation(q, lambda_s):  # :30-38
    return q < 1.2321 - .124616 * math.log(lambda_s)


NEG_INF = -math.inf


def _out(x, b):
    return x < b[0] or x > b[1]


def log_prior_D(pos, PD):  # logPriorStandard_D::eval :409-436
    chirp = math.exp(pos[7])
    eta = pos[8]
    if eta < .0 or eta > .25:
        return NEG_INF
    m1, m2 = calculate_mass1(chirp, eta), calculate_mass2(chirp, eta)
    if _out(m1, PD["mass1_prior"]) or _out(m2, PD["mass2_prior"]):
        return NEG_INF
    if _out(pos[0], PD["RA_bounds"]) or _out(pos[1], PD["sinDEC_bounds"]):
        return NEG_INF
    if pos[2] < 0 or pos[2] > math.pi or pos[3] < -1 or pos[3] > 1 or pos[4] < 0 or pos[4] > 2 * math.pi:
        return NEG_INF
    if pos[5] < PD["T_merger"] - .1 or pos[5] > PD["T_merger"] + .1:
        return NEG_INF
    if _out(math.exp(pos[6]), PD["DL_prior"]):
        return NEG_INF
    if _out(pos[9], PD["spin1_prior"]) or _out(pos[10], PD["spin2_prior"]):
        return NEG_INF
    return (math.log(aligned_spin_prior(pos[9])) + math.log(aligned_spin_prior(pos[10])) + math.log(chirpmass_eta_jac(chirp, eta))
            + 3 * pos[6])


def log_prior_P(pos, PD):  # logPriorStandard_P::eval :438-469
    chirp = math.exp(pos[7])
    eta = pos[8]
    if eta < .0 or eta > .25:
        return NEG_INF
    m1, m2 = calculate_mass1(chirp, eta), calculate_mass2(chirp, eta)
    if _out(m1, PD["mass1_prior"]) or _out(m2, PD["mass2_prior"]):
        return NEG_INF
    if _out(pos[0], PD["RA_bounds"]) or _out(pos[1], PD["sinDEC_bounds"]):
        return NEG_INF
    if pos[2] < 0 or pos[2] > math.pi or pos[3] < -1 or pos[3] > 1 or pos[4] < 0 or pos[4] > 2 * math.pi:
        return NEG_INF
    if pos[5] < PD["T_merger"] - .1 or pos[5] > PD["T_merger"] + .1:
        return NEG_INF
    if _out(math.exp(pos[6]), PD["DL_prior"]):
        return NEG_INF
    for i, name in ((9, "a1_prior"), (10, "a2_prior"), (11, "ctheta1_prior"), (12, "ctheta2_prior"), (13, "phi1_prior"), (14, "phi2_prior")):
        if _out(pos[i], PD[name]):
            return NEG_INF
    return math.log(chirpmass_eta_jac(chirp, eta)) + 3 * pos[6]


def log_prior_NRT(pos, PD, pv2):  # logPriorStandard_D_NRT :383-407, logPriorStandard_P_NRT :485-509
    chirp = math.exp(pos[7])
    m1, m2 = calculate_mass1(chirp, pos[8]), calculate_mass2(chirp, pos[8])
    q = m2 / m1
    t0 = 15 if pv2 else 11
    factor = 0
    if PD["tidal_love"]:
        if _out(math.exp(pos[t0]), PD["tidal_s_prior"]):
            return NEG_INF
        if tidal_love_boundary_violation(q, math.exp(pos[11])):  # pos[11] in both variants (:499)
            return NEG_INF
        factor += pos[11]
    else:
        if _out(math.exp(pos[t0]), PD["tidal1_prior"]) or _out(math.exp(pos[t0 + 1]), PD["tidal2_prior"]):
            return NEG_INF
        factor += pos[t0]
        factor += pos[t0 + 1]
    return (log_prior_P(pos, PD) if pv2 else log_prior_D(pos, PD)) + factor


def standard_log_prior(pos, PD, pv2, nrt):
    """[_mod] -> [_NRT] -> base (:321-335, 337-353, 471-483, 511-526)."""
    base = 15 if pv2 else 11
    if nrt:
        base += 1 if PD["tidal_love"] else 2
    for i in range(base, len(pos)):
        if _out(pos[i], PD["mod_priors"][i - base]):
            return NEG_INF
    if nrt:
        return log_prior_NRT(pos, PD, pv2)
    return log_prior_P(pos, PD) if pv2 else log_prior_D(pos, PD)


# ---- Fisher post-processing: src/mcmc_gw.cpp:2136-2189, src/mcmc_sampler_internals.cpp:643-713 ------------------------------

def fisher_transformations(F, pv2, alpha_unit_fix=False, ppE_Nmod=0, param=None):
    F = np.array(F, dtype=np.float64)
    pi2 = 4 * math.pi * math.pi
    for i, v in ((0, 1. / pi2), (1, 1. / 4), (2, 1. / pi2), (3, 1. / 4), (4, 1. / pi2), (5, 1. / .01), (8, 1. / .25), (9, 1. / 4), (10, 1. / 4)):
        F[i, i] += v
    if pv2:
        for i, v in ((11, 1. / 4), (12, 1. / 4), (13, 1. / pi2), (14, 1. / pi2)):
            F[i, i] += v
    if alpha_unit_fix:
        dim = F.shape[0]
        base = dim - ppE_Nmod
        factor = 4 * param[base] ** (3. / 4.) * 1000 / 299792458.
        F[base, :] *= factor
        F[:, base] *= factor
    return F


def eigen_system(F):
    """Eigen::SelfAdjointEigenSolver: ascending eigenvalues; row i of the result = eigenvector i (fisher_vecs layout, :687)."""
    vals, vecs = np.linalg.eigh(np.asarray(F))
    return vals, vecs.T.copy()


# ---- the sampler -------------------------------------------------------------------------------------------------------------

def step_boundaries(T, fisher_exist, de_primed):  # assign_probabilities, non-RJ (:1200-1250, 1357-1362)
    p = [0., 0., 0., 0.]
    if not fisher_exist:
        p[0] = 1.
    elif not de_primed:
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
    return [b0, b1, b2, b3]


class Sampler:
    """One object = the reference's `sampler` struct for a fixed-ladder, non-RJ PTMCMC run."""

    def __init__(self, loglike, log_prior, temps, initial, seed, swp_freq=5, swap_rate=None, history_length=1000, history_update=10,
                 fisher=None, fisher_update_number=200, check_stepsize_freq=50):
        self.ll_fn, self.lp_fn, self.fisher_fn = loglike, log_prior, fisher
        self.T = [float(t) for t in temps]
        self.C, self.P = np.shape(initial)
        self.pos = [list(map(float, row)) for row in initial]
        self.seed = seed
        self.swp_freq = swp_freq
        self.swap_rate = 1. / swp_freq if swap_rate is None else swap_rate
        self.H, self.history_update = history_length, history_update
        self.fisher_exist = fisher is not None
        self.fisher_update_number = fisher_update_number
        self.check_stepsize_freq = check_stepsize_freq
        self.step = 0
        self.sweep = 0
        C, P = self.C, self.P
        self.ll = [float(x) for x in self.ll_fn(np.array(self.pos))]
        self.lp = [self.lp_fn(p) for p in self.pos]
        self.widths = [[.05] * P + [1., .05, .5] for _ in range(C)]  # :1998-2004
        self.hist = [[None] * self.H for _ in range(C)]
        self.hist_pos = [0] * C
        for c in range(C):
            self.hist[c][0] = list(self.pos[c])
        self.fvals = [[1.] * P for _ in range(C)]
        self.fvecs = [np.eye(P) for _ in range(C)]
        self.fisher_ct = [fisher_update_number] * C  # :2011
        self.gauss_ct = [[[0, 0, 0, 0] for _ in range(P)] for _ in range(C)]
        self.type_last = [[0, 0, 0, 0] for _ in range(C)]
        self.ct = [dict(step=[0, 0], gauss=[0, 0], de=[0, 0], fisher=[0, 0], swap=[0, 0], fisher_updates=0) for _ in range(C)]
        self.pending = None
        if self.fisher_exist:
            # assign_initial_pos (:3182-3212): every chain's matrix at its initial position; update_fisher resets the counter
            for c in range(C):
                vals, vecs = self.fisher_fn(c, np.array(self.pos[c]))
                if not (np.isnan(vals).any() or np.isnan(vecs).any()):
                    self.fvals[c], self.fvecs[c] = list(vals), np.array(vecs)
                    self.ct[c]["fisher_updates"] += 1
                self.fisher_ct[c] = 0

    # -- proposals (one chain) --
    def _propose(self, c, s):
        P, T = self.P, self.T[c]
        primed = s > self.H
        bounds = step_boundaries(T, self.fisher_exist, primed)
        alpha, u_acc = uniform2(self.seed, s, c, DRAW_TYPE_ACCEPT)
        u_pick, u_pick2 = uniform2(self.seed, s, c, DRAW_PICK)
        z = normal_from(*uniform2(self.seed, s, c, DRAW_NORMAL))
        cur = self.pos[c]
        sel = 0
        if alpha < bounds[0]:  # gaussian_step :364-397
            kind = STEP_GAUSS
            sel = int(u_pick * P)
            prop = list(cur)
            prop[sel] = z * self.widths[c][sel] + cur[sel]
        elif alpha < bounds[1]:  # diff_ev_step :930-975
            kind = STEP_DE
            i = int(self.H * u_pick)
            j = (i + 1 + int((self.H - 1) * u_pick2)) % self.H
            beta, _ = uniform2(self.seed, s, c, DRAW_DE_SCALE)
            a = 1.
            if beta < .9:
                a = z * self.widths[c][P + 0]
            prop = [cur[k] + a * (self.hist[c][i][k] - self.hist[c][j][k]) for k in range(P)]
        else:  # fisher_step :424-518, 627
            kind = STEP_FISHER
            if self.fisher_ct[c] == self.fisher_update_number:
                vals, vecs = self.fisher_fn(c, np.array(cur))
                if not (np.isnan(vals).any() or np.isnan(vecs).any()):
                    self.fvals[c], self.fvecs[c] = list(vals), np.array(vecs)
                    self.ct[c]["fisher_updates"] += 1
                self.fisher_ct[c] = 0
            beta = int(P * u_pick)
            a = z * self.widths[c][P + 2]
            scaling = 10. if abs(self.fvals[c][beta]) < 10 else abs(self.fvals[c][beta]) / T
            sc = a / math.sqrt(scaling)
            prop = [cur[i] + sc * self.fvecs[c][beta][i] for i in range(P)]
            self.fisher_ct[c] += 1
        return kind, sel, prop, u_acc

    def _finish(self, c, s, kind, sel, prop, u_acc, proposed_ll, proposed_lp):
        # mcmc_step :84-146
        T = self.T[c]
        current_lp = self.lp[c]
        if current_lp == NEG_INF or proposed_lp == NEG_INF or math.isnan(proposed_ll):
            mh = NEG_INF
        else:
            mh = (-self.ll[c] + proposed_ll) / T - current_lp + proposed_lp
        acc = not (mh < math.log(u_acc)) if u_acc > 0 else True
        name = {STEP_GAUSS: "gauss", STEP_DE: "de", STEP_FISHER: "fisher"}[kind]
        if acc:
            self.pos[c], self.ll[c], self.lp[c] = list(prop), proposed_ll, proposed_lp
        self.ct[c]["step"][0 if acc else 1] += 1
        self.ct[c][name][0 if acc else 1] += 1
        if kind == STEP_GAUSS:
            self.gauss_ct[c][sel][0 if acc else 1] += 1
        # PTMCMC_MH_step_incremental :4603-4642
        chain_pos = s + 1
        primed = s > self.H
        if (not primed) or chain_pos % self.history_update == 0:  # update_history :2198-2219
            hp = self.hist_pos[c]
            hp = hp + 1 if hp < self.H - 1 else 0
            self.hist_pos[c] = hp
            self.hist[c][hp] = list(self.pos[c])
        if chain_pos % self.check_stepsize_freq == 0:  # update_step_widths :1623-1703
            b = step_boundaries(T, self.fisher_exist, primed)
            lo, hi = .2, .60 - .2 / T

            def tuned(w, a, r):
                if a + r == 0:
                    return w
                frac = a / (a + r)
                return w * .9 if frac < lo else (w * 1.1 if frac > hi else w)
            P = self.P
            if b[0] != 0:
                for i in range(P):
                    g = self.gauss_ct[c][i]
                    self.widths[c][i] = tuned(self.widths[c][i], g[0] - g[2], g[1] - g[3])
                    g[2], g[3] = g[0], g[1]
            tl = self.type_last[c]
            if b[1] - b[0] != 0:
                self.widths[c][P] = tuned(self.widths[c][P], self.ct[c]["de"][0] - tl[0], self.ct[c]["de"][1] - tl[1])
                tl[0], tl[1] = self.ct[c]["de"]
            if b[3] - b[2] != 0:
                self.widths[c][P + 2] = tuned(self.widths[c][P + 2], self.ct[c]["fisher"][0] - tl[2], self.ct[c]["fisher"][1] - tl[3])
                tl[2], tl[3] = self.ct[c]["fisher"]

    def _swap_sweep(self):
        # PTMCMC_MH_step_incremental :4646-4654, chain_swap :1086-1118, single_chain_swap :1121-1184
        gate, _ = uniform2(self.seed, self.sweep, 0, DRAW_SWAP_GATE)
        if gate < self.swap_rate:
            for i in range(self.C - 1):
                T1, T2 = self.T[i], self.T[i + 1]
                ok = False
                if T1 != T2:
                    ll1, ll2 = self.ll[i], self.ll[i + 1]
                    pw = (ll1 - ll2) / T2 - (ll1 - ll2) / T1
                    alpha, _ = uniform2(self.seed, self.sweep, i, DRAW_SWAP)
                    try:
                        ratio = math.exp(pw)
                    except OverflowError:
                        ratio = math.inf
                    ok = not (ratio < alpha)
                if ok:
                    self.pos[i], self.pos[i + 1] = self.pos[i + 1], self.pos[i]
                    self.ll[i], self.ll[i + 1] = self.ll[i + 1], self.ll[i]
                    self.lp[i], self.lp[i + 1] = self.lp[i + 1], self.lp[i]
                for k in (i, i + 1):
                    self.ct[k]["swap"][0 if ok else 1] += 1
        self.sweep += 1

    def run(self, n_steps):
        since = self.step % self.swp_freq
        for _ in range(n_steps):
            s = self.step
            props = [self._propose(c, s) for c in range(self.C)]
            lls = self.ll_fn(np.array([p[2] for p in props]))
            for c, (kind, sel, prop, u_acc) in enumerate(props):
                self._finish(c, s, kind, sel, prop, u_acc, float(lls[c]), self.lp_fn(prop))
            self.step += 1
            since += 1
            if since == self.swp_freq:
                self._swap_sweep()
                since = 0
