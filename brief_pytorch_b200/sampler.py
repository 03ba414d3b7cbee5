"""Samplers — same iterator protocol as the reference's main.py:38-163 (== utils/sampler.py:9-94):
`len(s) == max_steps`, `next(s) -> (coords, data, weight)`.

The fused fit kernel does this gather on chip (raw voxel -> normalise -> weight rule, coordinates from
axis tables), so these classes are only needed by callers that want the three tensors; they produce them
with the standalone gather kernel (brief_gather) from a bound SirenGroup instead of three advanced-index ops
over materialised fp32 copies.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .group import SirenGroup


class _GroupSampler:
    def __init__(self, group: SirenGroup, net: int, sample_count: int):
        self.group, self.net, self.sample_count = group, net, int(sample_count)
        spec = group.specs[net]
        self.pop_size = 1
        for n in spec.dims:
            self.pop_size *= int(n)
        self.index = 0
        self.last_idx: Optional[torch.Tensor] = None

    def __len__(self):
        return self.sample_count

    def __iter__(self):
        self.index = 0
        return self


class RandompointSampler(_GroupSampler):
    """`sample_size` voxel indices with replacement per step.  `generator='torch'` replays the reference's
    CPU torch.randint stream (parity runs); `generator='device'` uses the kernels' Philox stream."""

    def __init__(self, group: SirenGroup, net: int, sample_size: int, sample_count: int, generator: str = "torch",
                 seed: int = 42):
        super().__init__(group, net, sample_count)
        self.sample_size, self.generator, self.seed = int(sample_size), generator, seed

    def __next__(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if self.index >= self.sample_count:
            raise StopIteration
        if self.generator == "torch":
            idx = torch.randint(0, self.pop_size, (self.sample_size,)).to(self.group.device)
        else:
            from .group import sample_indices
            idx = sample_indices(self.seed, self.index, self.net, self.sample_size, self.pop_size, self.group.device)
        self.last_idx = idx
        self.index += 1
        return self.group.gather(self.net, idx)


class RandomCubeSampler(_GroupSampler):
    """The shipped configuration of the reference's cube sampler (cube_len clamped to the block, cube_count 1):
    every step is the whole block in voxel order, shaped [1, d, h, w, C] like the reference."""

    def __init__(self, group: SirenGroup, net: int, sample_count: int, cube_count: int = 1, cube_len=None):
        super().__init__(group, net, sample_count)
        dims = [int(n) for n in group.specs[net].dims]
        if cube_count != 1 or (cube_len is not None and any(int(c) < n for c, n in zip(cube_len, dims))):
            raise NotImplementedError("only whole-block cubes (the shipped cube_len/cube_count) are fused")
        self.dims = dims

    def __next__(self):
        if self.index >= self.sample_count:
            raise StopIteration
        self.index += 1
        c, d, w = self.group.gather(self.net, None, self.pop_size)
        shape = (1, *self.dims)
        return c.reshape(*shape, -1), d.reshape(*shape, 1), w.reshape(*shape, 1)
