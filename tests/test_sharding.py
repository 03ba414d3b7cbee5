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


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import zlib
        import yaml
        from brief_pytorch_b200 import synth
        from brief_pytorch_b200.CompressFramework import NFGR
        from brief_pytorch_b200.group import pack_module_params
        from test_framework import VESSEL_YAML
        o = yaml.safe_load(VESSEL_YAML)
        o["Compress"]["divide"]["divide_type"] = "total_1_2_4"
        o["Compress"]["param"]["filesize_ratio"] = 16
        o["Compress"]["checkpoints"] = "none"
        vol = synth.vessel((16, 48, 96), seed=7)
        blocks, mine = NFGR(o, rank, "f16", reproducible=True).compress_divide(vol, None, max_steps=12, rank=rank, world=world)
        owner = [0] * len(blocks)
        for i in mine:
            owner[i] = rank
        own = torch.tensor([float(i in mine) for i in range(len(blocks))], device="cuda")
        dist.all_reduce(own)
        assert (own == 1).all()  # every block fitted by exactly one rank
        crc = torch.tensor([[float(zlib.crc32(pack_module_params(blocks[i].module).tobytes()))] for i in mine],
                           dtype=torch.float64, device="cuda").reshape(-1, 1)
        who = torch.zeros(len(blocks), device="cuda")
        for i in mine:
            who[i] = rank
        dist.all_reduce(who)
        table = sharding.gather_block_stats(crc, [int(x) for x in who.tolist()])
        single = None
        if rank == 0:
            ref_blocks, _ = NFGR(o, 0, "f16", reproducible=True).compress_divide(vol, None, max_steps=12)
            single = [float(zlib.crc32(pack_module_params(b.module).tobytes())) for b in ref_blocks]
        q.put((rank, table[:, 0].cpu().tolist(), single))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_sharding_gives_bit_identical_blocks_nccl():
    """SURVEY 8(e) on hardware: the same volume fitted by 2 ranks (one process per GPU, NCCL) and by 1 rank gives
    bit-identical parameters for every block (reproducible slicing; the only collectives are the stats gathers)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, table, single = q.get(timeout=600)
        res[rank] = (table, single)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res[0][0] == res[1][0] == res[0][1]
