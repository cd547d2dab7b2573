"""
CPU: familiarity evaluation.  `pn_responses` and the `familiarity_par` indexing are checked
against the reference's own expressions (agent/agent.py:742-744, sim/simulation.py:3085-3090); the whitening and the
perfect memory are InvertPy's upstream (parity unpinned) and are checked against their stated definitions in plain
NumPy.  World-size-2 gloo run: every rank evaluates its block of positions, only the familiarity rows travel.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from invertsy_b200 import familiarity as fam


def _data(V=60, L=25, N=40, C=5, seed=0):
    rng = np.random.RandomState(seed)
    route = rng.uniform(-0.2, 1.3, (L, N, C)).astype(np.float32)
    views = rng.uniform(-0.2, 1.3, (V, N, C)).astype(np.float32)
    views[:L:3] = route[:L:3]                      # some views are exactly on the route
    return views, route


class _NumpyMemory:
    """stands in for the CUDA perfect memory in the CPU tests of the host logic (the definition, in NumPy float64)"""

    def store(self, v):
        self.db = torch.as_tensor(v).double().numpy()
        return self

    def familiarity(self, q):
        return torch.from_numpy(fam.reference_familiarity(torch.as_tensor(q).double().numpy(), self.db)[0])


def test_pn_responses_is_the_reference_expression():
    views, _ = _data()
    ref = np.clip(views.mean(axis=2), 0, 1)        # agent.py:742-744 applied to every view
    assert np.allclose(fam.pn_responses(views).numpy(), ref, atol=1e-7)


def test_whitening_and_perfect_memory_match_their_definitions():
    views, route = _data()
    pv, pr = np.clip(views.mean(2), 0, 1).astype(np.float64), np.clip(route.mean(2), 0, 1).astype(np.float64)
    # definition in NumPy
    m = pr.mean(0)
    d, e = np.linalg.eigh((pr - m).T @ (pr - m) / pr.shape[0])
    order = np.argsort(d)[::-1][:10]
    w = e[:, order] / np.sqrt(np.maximum(d[order], 0) + 1e-5)
    yv, yr = (pv - m) @ w, (pr - m) @ w
    ref = 1 - np.sqrt(((yv[:, None, :] - yr[None, :, :]) ** 2).mean(2)).min(1)
    wh = fam.Whitening(nb_output=10, dtype=torch.float64).fit(pr)
    got = fam.evaluate(views.astype(np.float64), route.astype(np.float64), whitening=wh, memory=_NumpyMemory()).numpy()
    # eigenvectors are defined up to a sign: distances are not affected
    assert np.allclose(got, ref, atol=1e-6)
    assert np.allclose(got[:25:3], 1.0, atol=1e-5)              # views on the route are fully familiar
    white = wh(pr).numpy()
    assert np.allclose(white.T @ white / pr.shape[0], np.eye(10), atol=1e-3)   # whitened: unit covariance
    f2, idx = fam.reference_familiarity(yv, yr)
    assert np.allclose(f2, ref, atol=1e-12) and np.array_equal(idx[:25:3], np.arange(0, 25, 3))


def test_the_cuda_memory_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mem = fam.PerfectMemory().store(np.zeros((3, 8), dtype=np.float32))
    with pytest.raises(Exception):
        mem.familiarity(np.zeros((2, 8), dtype=np.float32))


def test_familiarity_par_uses_the_reference_indexing():
    L, K, O = 6, 3, 4
    V = L * K * O
    f = np.arange(V, dtype=np.float64)
    out = fam.familiarity_par(f, L, K, O)
    for j in range(V):                              # simulation.py:3085-3090, verbatim
        m = (j // O) % L
        k = j // (L * O)
        ori = j % O
        assert out[m, k, ori] == f[j]
    assert fam.familiarity_map(f, L, K, O).shape == (L, K, O)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from invertsy_b200.views import shard_positions
        P, H = 7, 4
        views, route = _data(V=P * H, L=9)
        full = fam.evaluate(views, route, memory=_NumpyMemory())
        a, b = shard_positions(P, rank, world)
        got = fam.evaluate_sharded(views[a * H:b * H], route, P, H, memory=_NumpyMemory())
        q.put((rank, bool(torch.allclose(got, full))))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_gathers_the_rows_in_order():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
