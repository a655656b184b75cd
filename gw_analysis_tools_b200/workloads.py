"""Synthetic, seeded workloads of the shapes BASELINE.json names.

Pure input preparation (numpy): frequency grids, an analytic PSD, injected parameters and the walker scatter.  The same
arrays feed the CUDA path, the oracle and the CPU baseline, so every arm sees identical inputs.
Sampling-vector layout is the reference's "MCMC_" parameterisation (src/fisher.cpp:1909-1933, 1845-1877):
    RA, sin DEC, psi, cos iota, phiRef, tc, ln DL, ln Mc, eta, chi1, chi2                       (IMRPhenomD, 11)
    RA, sin DEC, psi, cos iota, phiRef, tc, ln DL, ln Mc, eta, a1, a2, cos t1, cos t2, ph1, ph2   (IMRPhenomPv2, 15)
followed by ln Lambda_s (IMRPhenomD_NRT with tidal_love) and/or the modification parameters.
"""
import dataclasses

import numpy as np

from . import abi

MSOL_SEC = 4.925491025543575903411922162094833998e-6
SEED0 = 20261017


def aligo_analytic_psd(f):
    """Square of the reference's analytic aLIGO amplitude spectral density (src/detector_util.cpp:288-295)."""
    x = 70.0 / np.asarray(f, dtype=np.float64)
    return 3e-48 * (x ** 4 + 2 + 2 * x * x) / 5


@dataclasses.dataclass
class Workload:
    name: str
    method: str
    detectors: list
    f: np.ndarray            # [L]
    psd: np.ndarray          # [D][L]
    T_segment: float
    gmst: float
    inj: np.ndarray          # injected sampling vector [P]
    params: np.ndarray       # walkers [W][P]
    mod: object = None       # abi.Mod or None
    data: np.ndarray = None  # [D][L] complex, filled by `with_injection`

    @property
    def W(self):
        return self.params.shape[0]

    @property
    def L(self):
        return self.f.size

    @property
    def D(self):
        return len(self.detectors)

    @property
    def P(self):
        return self.params.shape[1]


def _chirp_eta(m1, m2):
    return (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2, m1 * m2 / (m1 + m2) ** 2


def _base_vector(m1, m2, chi1, chi2, DL, T_segment):
    mc, eta = _chirp_eta(m1, m2)
    # RA, sin DEC, psi, cos iota, phiRef, tc, ln DL, ln Mc, eta, chi1, chi2
    return np.array([0.275, np.sin(-0.44), 0.2, np.cos(0.51), 2.0, 2.0, np.log(DL), np.log(mc), eta, chi1, chi2])


_SIGMA11 = np.array([0.05, 0.02, 0.05, 0.02, 0.05, 1e-3, 0.1, 1e-3, 5e-3, 0.05, 0.05])


def _scatter(rng, inj11, W):
    p = inj11[None, :] + rng.standard_normal((W, 11)) * _SIGMA11[None, :]
    p[:, 1] = np.clip(p[:, 1], -0.999, 0.999)
    p[:, 3] = np.clip(p[:, 3], -0.999, 0.999)
    p[:, 8] = np.clip(p[:, 8], 0.05, 0.2499)
    p[:, 9:11] = np.clip(p[:, 9:11], -0.9, 0.9)
    return p


def _grid(fmin, df, L):
    return fmin + df * np.arange(L, dtype=np.float64)


def make(config, W=None, L=None, masses=None, seed=None):
    """Build BASELINE config `config` (1..5; 3 = Fisher sources uses `fisher_sources`).  W/L override the sizes."""
    rng = np.random.default_rng(SEED0 + config if seed is None else seed)
    gmst = 2.1
    if config == 1:
        W, L = W or 1024, L or 8192
        m1, m2 = masses or (36.0, 29.0)
        f = _grid(20.0, 1.0 / 8, L)
        T = 8.0
        dets = ["Hanford", "Livingston"]
        inj = _base_vector(m1, m2, 0.3, 0.2, 500.0, T)
        params = _scatter(rng, inj, W)
        return Workload("cfg1_IMRPhenomD_2det", "IMRPhenomD", dets, f, np.tile(aligo_analytic_psd(f), (2, 1)), T, gmst,
                        inj, params)
    if config == 2:
        W, L = W or 4096, L or 16384
        m1, m2 = masses or (36.0, 29.0)
        f = _grid(20.0, 1.0 / 8, L)
        T = 8.0
        dets = ["Hanford", "Livingston", "Virgo"]
        b = _base_vector(m1, m2, 0.0, 0.0, 500.0, T)
        # a1, a2, cos tilt1, cos tilt2, phi1, phi2
        inj = np.concatenate([b[:9], [0.4, 0.3, 0.5, -0.2, 1.0, 2.5]])
        base = _scatter(rng, b, W)
        prec = np.column_stack([rng.uniform(0, 0.9, W), rng.uniform(0, 0.9, W), rng.uniform(-1, 1, W),
                                rng.uniform(-1, 1, W), rng.uniform(0, 2 * np.pi, W), rng.uniform(0, 2 * np.pi, W)])
        params = np.concatenate([base[:, :9], prec], axis=1)
        return Workload("cfg2_IMRPhenomPv2_3det", "IMRPhenomPv2", dets, f, np.tile(aligo_analytic_psd(f), (3, 1)), T,
                        gmst, inj, params)
    if config == 4:
        W, L = W or 16384, L or 16384
        m1, m2 = masses or (36.0, 29.0)
        f = _grid(20.0, 1.0 / 8, L)
        T = 8.0
        dets = ["Hanford", "Livingston", "Virgo"]
        b = _base_vector(m1, m2, 0.3, 0.2, 500.0, T)
        inj = np.concatenate([b, [10.0]])  # sqrt(alpha_dCS) in km
        params = np.concatenate([_scatter(rng, b, W), rng.uniform(0, 30, (W, 1))], axis=1)
        mod = abi.mod_defaults(ppE_Nmod=1, bppe=[-1.0])
        return Workload("cfg4_dCS_IMRPhenomD_3det", "dCS_IMRPhenomD", dets, f, np.tile(aligo_analytic_psd(f), (3, 1)), T,
                        gmst, inj, params, mod)
    if config == 5:
        W, L = W or 4096, L or (1 << 20)
        m1, m2 = masses or (1.5, 1.3)
        f = _grid(10.0, 1.0 / 256, L)
        T = 256.0
        dets = ["Hanford", "Livingston", "Virgo"]
        b = _base_vector(m1, m2, 0.02, 0.01, 100.0, T)
        inj = np.concatenate([b, [np.log(400.0)]])
        base = _scatter(rng, b, W)
        base[:, 9:11] = np.clip(base[:, 9:11], -0.05, 0.05)
        params = np.concatenate([base, rng.uniform(np.log(50.0), np.log(2000.0), (W, 1))], axis=1)
        mod = abi.mod_defaults(tidal_love=1, NSflag1=1, NSflag2=1)
        return Workload("cfg5_IMRPhenomD_NRT_3det", "IMRPhenomD_NRT", dets, f, np.tile(aligo_analytic_psd(f), (3, 1)), T,
                        gmst, inj, params, mod)
    raise ValueError("config must be 1, 2, 4 or 5 (3 = fisher_sources)")


def fisher_sources(S, seed=None):
    """Config 3: S random sources drawn like the reference's Fisher comparison (testing/fisher_comparison.cpp:76-106)."""
    rng = np.random.default_rng(SEED0 + 3 if seed is None else seed)
    out = []
    for _ in range(S):
        ma, mb = rng.uniform(3, 100, 2)
        m1, m2 = max(ma, mb), min(ma, mb)
        out.append(abi.source_defaults(
            mass1=m1, mass2=m2, Luminosity_Distance=rng.uniform(10, 1000),
            spin1=[0, 0, rng.uniform(-0.9, 0.9)], spin2=[0, 0, rng.uniform(-0.9, 0.9)],
            RA=rng.uniform(0, 2 * np.pi), DEC=np.arcsin(rng.uniform(-1, 1)), psi=rng.uniform(0, np.pi),
            incl_angle=np.arccos(rng.uniform(-1, 1)), phiRef=rng.uniform(0, 2 * np.pi), tc=rng.uniform(1, 6),
            gmst=2.1, f_ref=20.0, shift_time=1, shift_phase=1))
    return out
