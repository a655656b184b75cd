"""CPU tier for the batched PTMCMC step: the sampler's GWAT_HD mathematics, compiled as plain C++ in
tests/host_harness.cpp, against the restatement of the reference's rules in oracle/ptmcmc_ref.py; Philox against the
known-answer vectors of the Random123 distribution; the eigen-solver against LAPACK.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from gw_analysis_tools_b200 import abi
from oracle import ptmcmc_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


from gw_analysis_tools_b200.sampler import Prior  # noqa: E402  (ctypes layout only; nothing here touches the CUDA library)


PRIOR = dict(mass1_prior=[5, 80], mass2_prior=[3, 60], spin1_prior=[-.9, .9], spin2_prior=[-.9, .9], a1_prior=[0, .9], a2_prior=[0, .9],
             ctheta1_prior=[-1, 1], ctheta2_prior=[-1, 1], phi1_prior=[0, 2 * math.pi], phi2_prior=[0, 2 * math.pi],
             tidal1_prior=[1, 5000], tidal2_prior=[1, 5000], tidal_s_prior=[1, 5000], RA_bounds=[0, 2 * math.pi],
             sinDEC_bounds=[-1, 1], DL_prior=[10, 5000], T_merger=6.0, tidal_love=1,
             mod_priors=[[-5, 5]] * abi.MAX_MOD)


def c_prior(d):
    p = Prior()
    for k, v in d.items():
        if k == "mod_priors":
            for i, (lo, hi) in enumerate(v):
                p.mod_priors[i][0], p.mod_priors[i][1] = lo, hi
        elif isinstance(v, list):
            getattr(p, k)[0], getattr(p, k)[1] = v
        else:
            setattr(p, k, v)
    return p


@pytest.fixture(scope="module")
def hh():
    path = os.path.join(ROOT, "tests", "_build", "libgwat_host_harness.so")
    if not os.path.exists(path):
        pytest.fail("tests/_build/libgwat_host_harness.so missing: run __graft_entry__.build()")
    lib = C.CDLL(path)
    for n in ("hh_normal_from", "hh_log_prior", "hh_tuned_width"):
        getattr(lib, n).restype = C.c_double
    return lib


def test_philox_known_answers(hh):
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        assert ref.philox4x32_10(ctr, key) == want
        out = (C.c_uint * 4)()
        hh.hh_philox((C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), out)
        assert tuple(out) == want


def test_uniforms_and_normals_match_restatement(hh):
    rng = np.random.default_rng(5)
    us = []
    for _ in range(200):
        seed, step, chain, purpose = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**40)), int(rng.integers(0, 2**20)), int(rng.integers(0, 6))
        out = (C.c_double * 2)()
        hh.hh_uniform2(C.c_ulonglong(seed), C.c_ulonglong(step), chain, purpose, out)
        u = ref.uniform2(seed, step, chain, purpose)
        assert (out[0], out[1]) == u and 0 <= u[0] < 1 and 0 <= u[1] < 1
        us.append(u)
        z = hh.hh_normal_from(C.c_double(u[0]), C.c_double(u[1]))
        assert abs(z - ref.normal_from(*u)) <= 1e-14 * max(1, abs(z))
    us = np.array(us).ravel()
    assert abs(us.mean() - .5) < .05


def _positions(rng, n, pv2, nrt, nmod):
    """Sampling vectors scattered around and across the prior boundaries."""
    base = [rng.uniform(-.2, 6.5, n), rng.uniform(-1.1, 1.1, n), rng.uniform(-.2, 3.3, n), rng.uniform(-1.1, 1.1, n), rng.uniform(-.2, 6.5, n),
            rng.uniform(5.85, 6.15, n), np.log(rng.uniform(5, 6000, n)), np.log(rng.uniform(3, 60, n)), rng.uniform(-.01, .26, n)]
    if pv2:
        base += [rng.uniform(-.1, 1, n), rng.uniform(-.1, 1, n), rng.uniform(-1.1, 1.1, n), rng.uniform(-1.1, 1.1, n), rng.uniform(-.3, 6.5, n),
                 rng.uniform(-.3, 6.5, n)]
    else:
        base += [rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)]
    if nrt:
        base += [np.log(rng.uniform(.5, 6000, n)) for _ in range(nrt)]
    base += [rng.uniform(-6, 6, n) for _ in range(nmod)]
    return np.array(base).T.copy()


@pytest.mark.parametrize("pv2,nrt,nmod,tidal_love", [(0, 0, 0, 1), (1, 0, 0, 1), (0, 0, 2, 1), (1, 0, 1, 1), (0, 1, 0, 1), (0, 1, 1, 1), (1, 1, 0, 1), (0, 1, 0, 0), (1, 1, 0, 0)])
def test_standard_priors_match_restatement(hh, pv2, nrt, nmod, tidal_love):
    rng = np.random.default_rng(10 * pv2 + nrt + 100 * nmod)
    pd = dict(PRIOR, tidal_love=tidal_love)
    cp = c_prior(pd)
    pos = _positions(rng, 4000, pv2, nrt * (1 if tidal_love else 2), nmod)
    finite = 0
    for row in pos:
        want = ref.standard_log_prior(list(row), pd, bool(pv2), bool(nrt))
        got = hh.hh_log_prior(C.byref(cp), pv2, nrt, len(row), _p(row))
        if want == -math.inf:
            assert got == -math.inf
        else:
            finite += 1
            assert abs(got - want) <= 1e-13 * max(1, abs(want))
    if pv2 and nrt and tidal_love:
        # logPriorStandard_P_NRT tests the binary-Love boundary on pos[11] (cos tilt_1) instead of pos[15] (:499): with
        # q <= 1 < 1.2321 - .124616 pos[11] it rejects everything; the restatement and the kernel code keep that
        assert finite == 0
    else:
        assert 20 < finite < 3900  # the scatter exercises both outcomes


def test_step_probabilities(hh):
    for T in (1.0, 1.7, 12.0, 1e4):
        for fe in (0, 1):
            for pr in (0, 1):
                out = (C.c_double * 4)()
                hh.hh_step_boundaries(C.c_double(T), fe, pr, out)
                assert list(out) == ref.step_boundaries(T, bool(fe), bool(pr))
                assert abs(out[3] - 1) < 1e-15


def test_proposals_match_restatement(hh):
    rng = np.random.default_rng(3)
    P, H = 11, 50
    for _ in range(100):
        cur = rng.normal(size=P)
        widths = np.concatenate([rng.uniform(.01, .2, P), [1.3, .05, .7]])
        hist = rng.normal(size=(H, P))
        F = rng.normal(size=(P, P))
        vals, vecs = ref.eigen_system(F @ F.T * 1e3)
        u1, u2, beta, T = rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(1, 20)
        z = rng.normal()
        for kind in (ref.STEP_GAUSS, ref.STEP_DE, ref.STEP_FISHER):
            prop = np.empty(P)
            hh.hh_propose(kind, _p(cur), _p(prop), P, _p(widths), C.c_double(u1), C.c_double(u2), C.c_double(z), C.c_double(beta), H, _p(hist),
                          _p(vals), _p(np.ascontiguousarray(vecs)), C.c_double(T))
            if kind == ref.STEP_GAUSS:
                want = cur.copy()
                sel = int(u1 * P)
                want[sel] = z * widths[sel] + cur[sel]
            elif kind == ref.STEP_DE:
                i = int(H * u1)
                j = (i + 1 + int((H - 1) * u2)) % H
                assert i != j
                a = z * widths[P] if beta < .9 else 1.
                want = cur + a * (hist[i] - hist[j])
            else:
                b = int(P * u1)
                sc = 10. if abs(vals[b]) < 10 else abs(vals[b]) / T
                want = cur + (z * widths[P + 2]) / math.sqrt(sc) * vecs[b]
            assert np.allclose(prop, want, rtol=1e-15, atol=1e-15)


def test_acceptance_rules(hh):
    rng = np.random.default_rng(8)
    for _ in range(2000):
        cll, pll = rng.normal(1e4, 30, 2)
        clp, plp = rng.normal(0, 3, 2)
        T, u = rng.uniform(1, 50), rng.uniform()
        mh = (-cll + pll) / T - clp + plp
        assert hh.hh_mh_accept(C.c_double(cll), C.c_double(pll), C.c_double(clp), C.c_double(plp), C.c_double(T), C.c_double(u)) == int(not mh < math.log(u))
    inf = -math.inf
    assert hh.hh_mh_accept(C.c_double(1.), C.c_double(5.), C.c_double(0.), C.c_double(inf), C.c_double(1.), C.c_double(.5)) == 0
    assert hh.hh_mh_accept(C.c_double(1.), C.c_double(math.nan), C.c_double(0.), C.c_double(0.), C.c_double(1.), C.c_double(.5)) == 0
    for _ in range(2000):
        l1, l2 = rng.normal(1e4, 5, 2)
        T1, T2, a = rng.uniform(1, 5), rng.uniform(1, 5), rng.uniform()
        want = int(not math.exp((l1 - l2) / T2 - (l1 - l2) / T1) < a)
        assert hh.hh_swap_decision(C.c_double(l1), C.c_double(l2), C.c_double(T1), C.c_double(T2), C.c_double(a)) == want
    assert hh.hh_swap_decision(C.c_double(1.), C.c_double(2.), C.c_double(3.), C.c_double(3.), C.c_double(.1)) == -1
    assert hh.hh_tuned_width(C.c_double(1.), C.c_longlong(1), C.c_longlong(9), C.c_double(.2), C.c_double(.4)) == .9
    assert hh.hh_tuned_width(C.c_double(1.), C.c_longlong(9), C.c_longlong(1), C.c_double(.2), C.c_double(.4)) == 1.1
    assert hh.hh_tuned_width(C.c_double(1.), C.c_longlong(3), C.c_longlong(7), C.c_double(.2), C.c_double(.4)) == 1.
    assert hh.hh_tuned_width(C.c_double(1.), C.c_longlong(0), C.c_longlong(0), C.c_double(.2), C.c_double(.4)) == 1.


def test_jacobi_against_lapack(hh):
    rng = np.random.default_rng(4)
    for n in (11, 12, 15, 17):
        for trial in range(5):
            scale = 10.0 ** rng.uniform(-4, 6, n)  # graded like a Fisher matrix
            B = rng.normal(size=(n, n))
            A = (B @ B.T) * np.outer(scale, scale)
            vals, vecs = np.empty(n), np.empty((n, n))
            assert hh.hh_jacobi(_p(np.ascontiguousarray(A)), n, _p(vals), _p(vecs)) == 1
            w = np.linalg.eigvalsh(A)
            assert np.all(np.diff(vals) >= 0)
            assert np.allclose(vals, w, rtol=1e-8, atol=1e-10 * abs(w).max())
            assert np.allclose(vecs @ vecs.T, np.eye(n), atol=1e-12)
            assert np.allclose(vecs @ A @ vecs.T, np.diag(vals), atol=1e-9 * abs(w).max())
    nan = np.full((3, 3), np.nan)
    vals, vecs = np.empty(3), np.empty((3, 3))
    assert hh.hh_jacobi(_p(nan), 3, _p(vals), _p(vecs)) == 0


def test_fisher_transformations(hh):
    rng = np.random.default_rng(6)
    for pv2, dim, fix, nmod in ((0, 11, 0, 0), (1, 15, 0, 0), (0, 12, 1, 1)):
        F = rng.normal(size=(dim, dim))
        F = F + F.T
        param = rng.uniform(.1, 3, dim)
        got = F.copy()
        hh.hh_fisher_transformations(_p(got), dim, pv2, fix, nmod, _p(param))
        want = ref.fisher_transformations(F, bool(pv2), bool(fix), nmod, param)
        assert np.allclose(got, want, rtol=1e-15, atol=0)


def test_restated_sampler_samples_a_known_target():
    """The restatement itself: on a Gaussian 'likelihood' the cold chains of a tempered run recover mean and variance."""
    P, C_ = 11, 8
    mu = np.array([1., 0., 1.5, 0., 3., 6., math.log(400.), math.log(20.), .2, 0., 0.])
    sig = np.array([.2, .1, .2, .1, .3, .01, .1, .05, .01, .1, .1])

    def ll(x):
        return -0.5 * (((np.atleast_2d(x) - mu) / sig) ** 2).sum(axis=1)
    temps = np.array([1., 2., 4., 8.] * 2)
    s = ref.Sampler(ll, lambda p: 0.0, temps, np.tile(mu, (C_, 1)), seed=11, swp_freq=3, history_length=50, check_stepsize_freq=25)
    s.run(300)
    cold = []
    for _ in range(3000):
        s.run(1)
        cold += [s.pos[0], s.pos[4]]
    cold = np.array(cold)
    assert np.all(np.abs(cold.mean(axis=0) - mu) < 0.35 * sig)
    assert np.all(np.abs(cold.std(axis=0) / sig - 1) < 0.35)
    assert sum(c["swap"][0] for c in s.ct) > 0


def test_host_swap_sweep_matches_restated_sweep():
    """gwat_b200_swap_sweep_host (the threshold form the device sweep uses) against chain_swap/single_chain_swap as restated."""
    from gw_analysis_tools_b200 import sampler as smp
    rng = np.random.default_rng(12)
    for trial in range(20):
        C_ = int(rng.integers(2, 60))
        temps = np.tile(np.geomspace(1.0, 40.0, 5), 12)[:C_]
        if trial % 4 == 0:
            temps[C_ // 2:] = temps[C_ // 2]  # a run of equal temperatures: never swapped
        ll = rng.normal(1e4, 6, C_)
        seed, sweep = int(rng.integers(1, 2**40)), int(rng.integers(0, 1000))
        src, acc = smp.swap_sweep_host(ll, temps, seed, sweep)
        s = ref.Sampler.__new__(ref.Sampler)
        s.C, s.T, s.seed, s.sweep, s.swap_rate = C_, list(temps), seed, sweep, 2.0
        s.ll, s.lp, s.pos = list(ll), [0.0] * C_, [[float(i)] for i in range(C_)]
        s.ct = [dict(swap=[0, 0]) for _ in range(C_)]
        s._swap_sweep()
        assert [int(p[0]) for p in s.pos] == list(src)
        assert np.array_equal(np.array(s.ll), ll[src])
        n_acc = np.array([c["swap"][0] for c in s.ct])
        want = np.zeros(C_, dtype=int)
        want[:-1] += acc
        want[1:] += acc
        assert np.array_equal(n_acc, want)
    assert smp.draw_uniform2(5, 7, 3, 4) == ref.uniform2(5, 7, 3, 4)
