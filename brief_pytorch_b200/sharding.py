"""Multi-GPU sharding of the per-block networks.

Blocks are independent networks (main.py:484-532): the reference farms them out as one OS process each, placed on
the GPU with most free memory by a polling scheduler (main.py:573-579, utils/TasksManager.py:222-251) and exchanges
results through files.  Here one process per GPU owns a static, parameter-weighted share of the blocks; there is
NO collective on the fit path.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to gather
per-block loss statistics and decompressed blocks.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def fit_flops_per_sample(features: int, layers: int, coords_channel: int = 3, data_channel: int = 1) -> int:
    """6*P_w - 2*c*f: dW + dX for every layer, no dX for layer 0."""
    pw = coords_channel * features + (layers - 2) * features * features + features * data_channel
    return 6 * pw - 2 * coords_channel * features


def forward_flops_per_sample(features: int, layers: int, coords_channel: int = 3, data_channel: int = 1) -> int:
    return 2 * (coords_channel * features + (layers - 2) * features * features + features * data_channel)


def block_cost(features: int, layers: int, batch: int, steps: int, coords_channel: int = 3) -> float:
    """Parameter-weighted cost of fitting one block: steps * batch * fit FLOPs per sample."""
    return float(steps) * float(batch) * fit_flops_per_sample(features, layers, coords_channel)


def lpt_assign(costs: Sequence[float], n_ranks: int) -> List[int]:
    """Longest-processing-time-first: blocks in decreasing cost order (ties by index) go to the least-loaded
    rank (ties by rank).  Deterministic, so every rank computes the same owner table without communication."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    load = [0.0] * n_ranks
    owner = [0] * len(costs)
    for i in sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i)):
        r = min(range(n_ranks), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += float(costs[i])
    return owner


def my_blocks(owner: Sequence[int], rank: int) -> List[int]:
    return [i for i, r in enumerate(owner) if r == rank]


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_block_stats(local: torch.Tensor, owner: Sequence[int], group=None) -> torch.Tensor:
    """local: [n_local, k] fp32 rows (e.g. loss, mse, n_vox) for this rank's blocks in increasing block id.
    Returns the [n_blocks, k] table on every rank (all_gather on rows padded to the largest share)."""
    rank, world = _world()
    n_blocks, k = len(owner), int(local.shape[1])
    mine = my_blocks(owner, rank)
    assert local.shape[0] == len(mine), (local.shape, len(mine))
    if world == 1:
        return local.clone()
    cap = max(len(my_blocks(owner, r)) for r in range(world))
    pad = torch.zeros((cap, k), dtype=local.dtype, device=local.device)
    pad[:len(mine)] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    table = torch.zeros((n_blocks, k), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = my_blocks(owner, r)
        if ids:
            table[torch.tensor(ids, device=local.device)] = parts[r][:len(ids)]
    return table


def gather_blocks(local: Dict[int, torch.Tensor], owner: Sequence[int], shapes: Sequence[Sequence[int]], dst: int = 0,
                  group=None, dtype: Optional[torch.dtype] = None) -> Optional[Dict[int, torch.Tensor]]:
    """Bring every decompressed block to rank `dst` (merge_divided_data's input, utils/misc.py:430-445).  Every rank
    packs its blocks into ONE contiguous payload (a device copy at HBM rate) and the exchange is a single grouped
    send/recv — one message per sending rank instead of one per block and no pickled metadata: sizes follow from
    `owner`, `shapes` and `dtype` (inferred from the local blocks; pass it on a rank that owns none).  On `dst` the
    blocks are views into the per-sender receive buffers.  Returns {block id: tensor} on dst and None elsewhere.  For
    outputs that do not fit one GPU (the 4096^3 sweep) do not call this: keep the volume sharded by owner."""
    rank, world = _world()
    if world == 1:
        return dict(local)
    if dtype is None:
        if not local:
            raise ValueError("gather_blocks: pass dtype= on a rank that owns no block")
        dtype = next(iter(local.values())).dtype
    esz = int(torch.empty((), dtype=dtype).element_size())
    nbytes = []
    for sh in shapes:
        n = esz
        for x in sh:
            n *= int(x)
        nbytes.append(n)
    by_rank = [[b for b, r in enumerate(owner) if r == q] for q in range(world)]
    ref = next(iter(local.values())) if local else None
    dev = ref.device if ref is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    ops, keep, bufs = [], [], {}
    # payloads travel as raw bytes: NCCL has no 16-bit integer type (uint16 volumes are held as int16 bit patterns)
    if rank == dst:
        for q in range(world):
            if q != dst and by_rank[q]:
                bufs[q] = torch.empty(sum(nbytes[b] for b in by_rank[q]), dtype=torch.uint8, device=dev)
                ops.append(dist.P2POp(dist.irecv, bufs[q], q, group=group))
    elif by_rank[rank]:
        parts = [local[b].contiguous().view(torch.uint8).reshape(-1) for b in by_rank[rank]]
        payload = parts[0] if len(parts) == 1 else torch.cat(parts)
        keep.append(payload)
        ops.append(dist.P2POp(dist.isend, payload, dst, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    out: Dict[int, torch.Tensor] = {b: local[b] for b in by_rank[dst]}
    for q, buf in bufs.items():
        off = 0
        for b in by_rank[q]:
            out[b] = buf[off:off + nbytes[b]].view(dtype).reshape(tuple(int(x) for x in shapes[b]))
            off += nbytes[b]
    return out
