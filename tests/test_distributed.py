"""CPU, world_size 2, gloo: the multi-GPU host logic (position sharding, ordered gather of uneven shards)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(pos, yaw):
    """stands in for Renderer.render: a deterministic function of the pose, one row per view"""
    pos = torch.as_tensor(np.asarray(pos))
    yaw = torch.as_tensor(np.asarray(yaw))
    P, H = yaw.shape
    base = (pos[:, None, 0] * 1000 + pos[:, None, 1] * 10 + yaw).reshape(P * H, 1)
    return {"Y": base + torch.arange(5, dtype=base.dtype)[None, :], "hit": base.to(torch.int32).repeat(1, 3),
            "_keepalive": (pos, yaw)}


def _worker(rank, world, port, P, H, dst, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from invertsy_b200.distributed import render_sharded, shard
        rng = np.random.RandomState(0)
        pos = np.c_[rng.uniform(0, 10, P), rng.uniform(0, 10, P), np.full(P, 0.01)]
        yaw = rng.uniform(-3, 3, (P, H))
        full = _fake_render(pos, yaw)
        got, local, (a, b) = render_sharded(_fake_render, pos, yaw, dst=dst)
        _, _, blk = shard(pos, yaw)
        assert blk == (a, b) and local["Y"].shape[0] == (b - a) * H
        ok = True
        if dst is None or rank == dst:
            ok = all(torch.equal(got[k], full[k]) for k in ("Y", "hit"))
        else:
            ok = all(v is None for v in got.values())
        # a reduction before the gather: one number per view (heat-map style)
        red, _, _ = render_sharded(_fake_render, pos, yaw, reduce_fn=lambda o: {"m": o["Y"].mean(dim=1)})
        ok = ok and torch.allclose(red["m"], full["Y"].mean(dim=1))
        q.put((rank, bool(ok), (a, b)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,H,dst", [(7, 3, None), (1, 4, 0)])
def test_sharded_render_gathers_in_position_order(P, H, dst):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, H, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    blocks = [b for _, _, b in res]
    assert blocks[0][0] == 0 and blocks[-1][1] == P and blocks[0][1] == blocks[1][0]


def test_single_process_is_a_passthrough():
    from invertsy_b200.distributed import render_sharded
    pos = np.zeros((3, 3))
    yaw = np.zeros((3, 2))
    got, local, blk = render_sharded(_fake_render, pos, yaw)
    assert blk == (0, 3) and torch.equal(got["Y"], local["Y"])


def _worker_interleaved(rank, world, port, P, H, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from invertsy_b200.distributed import all_gather_interleaved
        from invertsy_b200.views import shard_positions_interleaved
        full = torch.arange(P * H, dtype=torch.float32) * 0.5 + 1.0          # value of view p * H + h
        idx = shard_positions_interleaved(P, rank, world)
        per = (P + world - 1) // world
        local = torch.zeros(per * H)
        local[:len(idx) * H] = full.view(P, H)[torch.as_tensor(idx)].reshape(-1)
        got = all_gather_interleaved(local, P, H)
        q.put((rank, bool(torch.equal(got, full)), len(idx)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,H", [(7, 3), (8, 2), (1, 4)])
def test_interleaved_shards_gather_back_into_view_order(P, H):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_interleaved, args=(r, world, port, P, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and sum(n for _, _, n in res) == P
