"""TEST INFRASTRUCTURE: the yardstick for finite-difference Fisher matrices.

The reference differentiates the detector response with an absolute step eps = 1e-8 (src/fisher.cpp:361-384), which
amplifies the rounding of every waveform evaluation by 1e8: its matrices are reproducible only to a noise level that depends
on the source (few cycles in band, weakly measured spins -> 1e-4 of sqrt(F_ii F_jj); a light, loud system -> 1e-7).  A
single other evaluation of the same mathematics shows that level: here the reference is run again on inputs moved by parts
in 1e14 -- six orders of magnitude below the stencil step, so the matrix it *should* produce is the same to 1e-6 of the
noise -- and, when the FMA-contracted build of the same sources exists (`make -C oracle noise`), once more with that.  The
largest normalised difference among these runs is the source's self-difference; an implementation agrees with the reference
when its own difference stays within FACTOR x that (and within BASELINE.json's 1e-6 where the reference is quieter).
"""
import ctypes as C

import numpy as np

from gw_analysis_tools_b200 import abi

FACTOR = 3.0
PERTURBED_RUNS = 4
REL = 1e-14


def normalised_error(F, R):
    """|dF_ij| / sqrt(R_ii R_jj) for one matrix or a stack of matrices."""
    F, R = np.asarray(F), np.asarray(R)
    dg = np.sqrt(np.abs(np.diagonal(R, axis1=-2, axis2=-1)))
    return np.abs(F - R) / (dg[..., :, None] * dg[..., None, :])


def perturbed(src, rng, rel=REL):
    t = abi.Source()
    C.memmove(C.addressof(t), C.addressof(src), C.sizeof(src))
    for name in ("mass1", "mass2", "Luminosity_Distance", "RA", "DEC", "psi", "incl_angle", "phiRef", "tc"):
        setattr(t, name, getattr(t, name) * (1 + rel * rng.uniform(-1, 1)))
    for k in range(3):
        t.spin1[k] *= 1 + rel * rng.uniform(-1, 1)
        t.spin2[k] *= 1 + rel * rng.uniform(-1, 1)
    return t


def reference_self_difference(oracle, method, srcs, dets, f, psd, dim, order, detector_index=-1, reference_index=0, nthreads=0,
                              runs=PERTURBED_RUNS, seed=1, with_fma=True, with_median=False):
    """Per source: max_ij normalised difference between the reference and `runs` re-evaluations of itself (see module doc);
    with_median: also the largest median_ij among those re-evaluations."""
    import os
    srcs = list(srcs)
    kw = dict(order=order, detector_index=detector_index, reference_index=reference_index, nthreads=nthreads)
    base = oracle.fisher_numerical_batch(method, srcs, dets, f, psd, dim, **kw)
    rng = np.random.default_rng(seed)
    worst = np.zeros(len(srcs))
    worst_median = np.zeros(len(srcs))
    others = [oracle.fisher_numerical_batch(method, [perturbed(s, rng) for s in srcs], dets, f, psd, dim, **kw) for _ in range(runs)]
    if with_fma and os.path.exists(oracle.LIB_PATH_FMA):
        others.append(oracle.fisher_numerical_batch(method, srcs, dets, f, psd, dim, fma_build=True, **kw))
    for o in others:
        e = normalised_error(o, base).reshape(len(srcs), -1)
        e = np.where(np.isfinite(e), e, 0.0)
        worst = np.fmax(worst, e.max(axis=1))
        worst_median = np.fmax(worst_median, np.median(e, axis=1))
    return (worst, worst_median) if with_median else worst
