"""Coordinate grids — same call signatures as the reference's utils/dataset.py:11-62.

The fused kernels never materialise the [N, ndim] grid: they look coordinates up in per-axis tables
(`axis_table`), which is what `create_coords` is built from here as well.
"""
from typing import Sequence, Tuple

import torch


def parse_coords_mode(mode: str) -> Tuple[float, float]:
    """'n11' -> (-1,1), '0p1' -> (0,1), otherwise 'lo,hi' (utils/dataset.py:12-20)."""
    named = {"n11": (-1.0, 1.0), "0p1": (0.0, 1.0)}
    if mode in named:
        return named[mode]
    lo, hi = (float(tok) for tok in mode.split(","))
    return lo, hi


def axis_table(n: int, mode: str = "n11") -> torch.Tensor:
    """torch.linspace on the CPU, exactly what the reference feeds to meshgrid for one axis."""
    lo, hi = parse_coords_mode(mode)
    return torch.linspace(lo, hi, int(n))


def create_coords(coords_shape: Sequence[int], mode: str = "n11") -> torch.Tensor:
    """[*coords_shape, ndim] grid with channel order = axis order (ij indexing)."""
    if len(coords_shape) not in (2, 3):
        raise NotImplementedError
    tables = [axis_table(n, mode) for n in coords_shape]
    mesh = torch.meshgrid(*tables, indexing="ij")
    return torch.stack(mesh, dim=-1)


def create_flattened_coords(coords_shape: Sequence[int], mode: str = "n11") -> torch.Tensor:
    grid = create_coords(coords_shape, mode)
    return grid.reshape(-1, grid.shape[-1])
