"""The one-chain-per-call queue (gwat_b200_queue_*, SURVEY 8b row 1: the sampler's likelihood callback).

CPU tier: argument checking without a context.  GPU tier: many threads call the callback the way the reference's pool
workers do; every value equals the value the batched call gives for the same vector (bit for bit: grouping must not change
a result), the golden values hold, and the calls really were merged into fewer launches.
"""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from gw_analysis_tools_b200 import abi, engine, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
LL_TOL = 1e-9  # BASELINE.json: log-likelihood relative error


def test_queue_create_rejects_bad_arguments():
    lib = engine.load_library()
    q = C.c_void_p()
    bad = lib.gwat_b200_queue_create(C.byref(q), None, b"IMRPhenomD", None, 11, C.c_double(0.0), C.c_double(8.0), 64, 8, C.c_double(100.0))
    assert bad == abi.ERR_ARG and not q
    nan = lib.gwat_b200_queue_loglike(None, None, None)
    assert np.isnan(nan)
    assert lib.gwat_b200_queue_stats(None, None, None, None) == abi.ERR_ARG
    lib.gwat_b200_queue_destroy(None)  # harmless, like free(NULL)


@pytest.fixture(scope="module")
def gold_mcmc():
    return np.load(os.path.join(GOLD, "mcmc_v1.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,threads", [(1, 16), (2, 7), (4, 16)])
def test_pool_of_threads_gets_batch_values(ctx, gold_mcmc, cfg, threads):
    wl = workloads.make(cfg, W=16, L=1024)
    ctx.set_network(wl.detectors, wl.f, wl.psd, gold_mcmc["cfg%d/data" % cfg])
    want = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    ref = gold_mcmc["cfg%d/logL" % cfg]
    assert (np.abs(want - ref) / np.abs(ref)).max() <= LL_TOL
    rounds = 12
    q = engine.LikelihoodQueue(ctx, wl.method, wl.params.shape[1], wl.gmst, wl.T_segment, wl.mod, max_batch=64,
                               expected_callers=threads, max_wait_us=20000.0)
    got = np.full((rounds, wl.W), np.nan)
    nxt = [0]
    lock = threading.Lock()

    def worker():
        while True:
            with lock:
                k = nxt[0]
                nxt[0] += 1
            if k >= rounds * wl.W:
                return
            got[k // wl.W, k % wl.W] = q.loglike(wl.params[k % wl.W])
    pool = [threading.Thread(target=worker) for _ in range(threads)]
    for t in pool:
        t.start()
    for t in pool:
        t.join()
    for r in range(rounds):
        assert np.array_equal(got[r], want)
    calls, batches, largest = q.stats()
    assert calls == rounds * wl.W
    assert batches < calls and largest > 1, (calls, batches, largest)
    q.close()


@pytest.mark.gpu
def test_single_caller_times_out_and_unphysical_point_is_nan(ctx, gold_mcmc):
    wl = workloads.make(1, W=16, L=1024)
    ctx.set_network(wl.detectors, wl.f, wl.psd, gold_mcmc["cfg1/data"])
    want = ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
    q = engine.LikelihoodQueue(ctx, wl.method, wl.params.shape[1], wl.gmst, wl.T_segment, wl.mod, max_batch=8, expected_callers=8,
                               max_wait_us=50.0)
    assert q.loglike(wl.params[3]) == want[3]  # nobody else joins: the leader gives up waiting after max_wait_us
    bad = wl.params[0].copy()
    bad[8] = 0.3  # eta > 1/4
    assert np.isnan(q.loglike(bad))
    assert q.stats() == (2, 2, 1)
    q.close()


@pytest.mark.gpu
def test_reference_style_thread_pool_through_the_cxx_layer(gold_mcmc):
    """tests/cxx/adapter_shim.cpp: std::thread workers call a std::function bound to gwat_b200::CallbackQueue."""
    path = os.path.join(ROOT, "tests", "_build", "libgwat_cxx_adapter.so")
    if not os.path.exists(path):
        pytest.fail("tests/_build/libgwat_cxx_adapter.so missing: run __graft_entry__.build()")
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    for cfg in (2, 4):
        wl = workloads.make(cfg, W=16, L=1024)
        data = gold_mcmc["cfg%d/data" % cfg]
        f = np.ascontiguousarray(wl.f)
        psd = np.ascontiguousarray(wl.psd)
        dre, dim = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
        W = 16 * 8
        params = np.ascontiguousarray(np.tile(wl.params, (8, 1)))
        out = np.full(W, np.nan)
        stats = (C.c_longlong * 3)()
        ok = lib.cxa_pool_loglike(wl.method.encode(), C.byref(wl.mod) if wl.mod is not None else None, params.shape[1], W,
                                  params.ctypes.data_as(dp), C.c_double(wl.gmst), C.c_double(wl.T_segment), ",".join(wl.detectors).encode(),
                                  f.ctypes.data_as(dp), psd.ctypes.data_as(dp), dre.ctypes.data_as(dp), dim.ctypes.data_as(dp), f.size, 8,
                                  out.ctypes.data_as(dp), stats)
        assert ok == 1
        ref = np.tile(gold_mcmc["cfg%d/logL" % cfg], 8)
        assert (np.abs(out - ref) / np.abs(ref)).max() <= LL_TOL
        assert stats[0] == W and stats[1] < W and stats[2] > 1, list(stats)
