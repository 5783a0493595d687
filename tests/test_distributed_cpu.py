"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: the batch partition and the one
collective of the design, the rollout-buffer all-gather.  The step path itself has no
collective and no CPU implementation, so nothing here computes a game step."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hironaka_b200.engine import gather_rollout, shard_range
        lo, hi = shard_range(total, rank, world)
        n = total // world  # equal shards for the gather (the engine gathers equal-shaped buffers)
        g = torch.Generator().manual_seed(100 + rank)
        obs = torch.rand((n, 63), generator=g)
        policy = torch.rand((n, 4), generator=g)
        value = torch.rand((n,), generator=g)
        G_obs, G_pol, G_val = gather_rollout([obs, policy, value])
        ok = G_obs.shape == (world, n, 63) and G_pol.shape == (world, n, 4) and G_val.shape == (world, n)
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            ok = ok and torch.equal(G_obs[r], torch.rand((n, 63), generator=gr))
            ok = ok and torch.equal(G_pol[r], torch.rand((n, 4), generator=gr))
            ok = ok and torch.equal(G_val[r], torch.rand((n,), generator=gr))
        # every game is owned by exactly one rank
        owned = torch.zeros(total, dtype=torch.int32)
        owned[lo:hi] = 1
        dist.all_reduce(owned)
        ok = ok and bool((owned == 1).all())
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == float(world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_shard_and_gather():
    world, total = 2, 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_gather_without_process_group():
    from hironaka_b200.engine import gather_rollout
    a, b = torch.arange(6.).reshape(3, 2), torch.arange(3.)
    ga, gb = gather_rollout([a, b])
    assert ga.shape == (1, 3, 2) and gb.shape == (1, 3) and torch.equal(ga[0], a)
