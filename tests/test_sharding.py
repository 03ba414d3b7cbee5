"""Multi-GPU host logic on CPU: deterministic LPT assignment, and the stats / block gathers over a world_size-2
gloo process group (the -m gpu runs use the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brief_pytorch_b200 import sharding


def test_lpt_assignment_is_balanced_and_deterministic():
    costs = [sharding.block_cost(f, 7, 100000, 80000) for f in (113, 56, 56, 39, 39, 39, 19, 19, 113, 228)]
    owner = sharding.lpt_assign(costs, 4)
    assert owner == sharding.lpt_assign(list(costs), 4)
    load = [sum(c for c, r in zip(costs, owner) if r == k) for k in range(4)]
    assert max(load) <= max(max(costs), sum(costs) / 4 * 4 / 3 + 1)  # LPT bound, or a single dominating block
    assert sorted(sum((sharding.my_blocks(owner, r) for r in range(4)), [])) == list(range(len(costs)))
    assert sharding.lpt_assign([1.0] * 8, 8) == list(range(8))
    assert sharding.lpt_assign([], 3) == []
    assert sharding.fit_flops_per_sample(56, 7) == 95088 and sharding.forward_flops_per_sample(56, 7) == 31808
    assert sharding.fit_flops_per_sample(22, 5) == 9108 and sharding.forward_flops_per_sample(13, 7) == 1794


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [5.0, 1.0, 4.0, 2.0, 3.0]
        owner = sharding.lpt_assign(costs, world)
        mine = sharding.my_blocks(owner, rank)
        local = torch.tensor([[float(b), 10.0 * b, float(rank)] for b in mine], dtype=torch.float32).reshape(-1, 3)
        table = sharding.gather_block_stats(local, owner)
        shapes = [(2, 3, b + 1) for b in range(len(costs))]
        blocks = {b: torch.full(shapes[b], 1000 + b, dtype=torch.int16) for b in mine}
        got = sharding.gather_blocks(blocks, owner, shapes, dst=0)
        q.put((rank, owner, table.numpy(), None if got is None else {b: t.numpy() for b, t in got.items()}))
    finally:
        dist.destroy_process_group()


def test_gathers_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, owner, table, blocks = q.get(timeout=120)
        res[rank] = (owner, table, blocks)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    owner = res[0][0]
    assert owner == res[1][0] and set(owner) == {0, 1}
    for rank in (0, 1):
        table = res[rank][1]
        np.testing.assert_array_equal(table[:, 0], np.arange(5))
        np.testing.assert_array_equal(table[:, 1], 10.0 * np.arange(5))
        np.testing.assert_array_equal(table[:, 2], np.float32(owner))
    assert res[1][2] is None
    blocks = res[0][2]
    assert sorted(blocks) == list(range(5))
    for b, arr in blocks.items():
        assert arr.shape == (2, 3, b + 1) and (arr == 1000 + b).all()
