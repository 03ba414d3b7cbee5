"""Samplers — same iterator protocol as the reference's main.py:38-163 (== utils/sampler.py:9-94):
`len(s) == max_steps`, `next(s) -> (coords, data, weight)`.

The fused fit kernel does this gather on chip (raw voxel -> normalise -> weight rule, coordinates from
axis tables), so these classes are only needed by callers that want the three tensors; they produce them
with the standalone gather kernel (brief_gather) from a bound SirenGroup instead of three advanced-index ops
over materialised fp32 copies.  Both samplers of the reference are complete here: random points, and sliding cubes of any
cube_len / cube_count (the whole-block form of the shipped configs included).
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
    """main.py:38-125: every step draws `cube_count` windows of `cube_len` voxels (clamped to the block) with
    replacement from all stride-1 window positions and returns them shaped [cube_count, *cube_len, C] like the
    reference.  With the shipped configuration (cube_len >= block, cube_count 1) the population is one cube and every
    step is the whole block in voxel order.  Constructing the sampler puts the network into this sampling mode
    (SirenGroup.set_cube_sampler), so `group.fit_step(sampler.last_idx)` fits exactly the samples it returned.
    `generator='torch'` replays the reference's CPU torch.randint draws; `generator='device'` uses the kernels' Philox
    stream (what SirenGroup.fit_run draws for this network)."""

    def __init__(self, group: SirenGroup, net: int, sample_count: int, cube_count: int = 1, cube_len=None,
                 generator: str = "torch", seed: int = 42):
        super().__init__(group, net, sample_count)
        dims = [int(n) for n in group.specs[net].dims]
        cube_len = [10000000] * len(dims) if cube_len is None else cube_len
        self.cube_len = [min(int(c), n) for c, n in zip(cube_len, dims)]  # main.py:49-50 (2-D: cube_len[0:2], :81-82)
        self.cube_count, self.generator, self.seed = int(cube_count), generator, seed
        self.pop_size = 1
        for c, n in zip(self.cube_len, dims):
            self.pop_size *= n - c + 1
        self.whole_block = self.pop_size == 1 and self.cube_count == 1
        group.set_cube_sampler(net, self.cube_count, self.cube_len)

    def __next__(self):
        if self.index >= self.sample_count:
            raise StopIteration
        if self.whole_block:
            idx, n = None, 1
            for c in self.cube_len:
                n *= c
            if self.generator == "torch":
                torch.randint(0, 1, (1,))  # the reference draws its one window every step (main.py:114): same RNG position
        elif self.generator == "torch":
            idx, n = self.group.cube_indices(self.net, torch.randint(0, self.pop_size, (self.cube_count,))), None
        else:
            idx, n = self.group.cube_indices(self.net, None, self.seed, self.index), None
        self.last_idx = idx
        self.index += 1
        c, d, w = self.group.gather(self.net, idx, n)
        shape = (self.cube_count, *self.cube_len)
        return c.reshape(*shape, -1), d.reshape(*shape, 1), w.reshape(*shape, 1)
