"""Synthetic uint16 volumes generated ON THE DEVICE, block by block, for the named shapes of BASELINE.json that do not
fit a host-side numpy generator (1024^3, 2048^3: `synth.py` would need tens of GB of float32 temporaries).  A block is a
function of its GLOBAL voxel coordinates and of (seed, block index) only, so every rank can produce exactly the blocks it
owns and a block's voxels do not depend on the sharding.  torch ops on CUDA tensors — data plumbing for the benchmark,
not part of the timed path.  Same structures as synth.py: `neuron` = sparse bright filaments + soma on a background
below 10000 (both classes of the value_10001_65535_0.1 weight rule), `hipct` = band-limited texture with a variance
contrast across the volume (non-uniform by_var budgets), `vessel` = dark background with bright tubes."""
from __future__ import annotations

from typing import Sequence

import torch


def _axes(vol_shape, lo, hi, device):
    # global coordinate of voxel i on an axis of n points: linspace(-1, 1, n)[i]
    out = []
    for k in range(3):
        n = int(vol_shape[k])
        i = torch.arange(int(lo[k]), int(hi[k]), device=device, dtype=torch.float32)
        out.append((i * (2.0 / max(n - 1, 1)) - 1.0))
    z, y, x = out
    return z[:, None, None], y[None, :, None], x[None, None, :]


def _noise(shape, std, seed, block_id, device):
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 1000003 + int(block_id)) & 0x7FFFFFFFFFFF)
    return torch.randn(tuple(shape), generator=g, device=device, dtype=torch.float32) * std


def _params(n, k, seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return (torch.rand((n, k), generator=g) * 2 - 1).tolist()


def block(kind: str, vol_shape: Sequence[int], lo: Sequence[int], hi: Sequence[int], seed: int, block_id: int,
          device) -> torch.Tensor:
    """uint16 voxels (as int16 bit patterns, contiguous [d,h,w]) of the block [lo, hi) of a `kind` volume."""
    z, y, x = _axes(vol_shape, lo, hi, device)
    shape = tuple(int(h - l) for l, h in zip(lo, hi))
    if kind == "neuron":
        vol = 1500.0 + _noise(shape, 300.0, seed, block_id, device)
        for a, b, c, d in _params(24, 4, seed):
            dist2 = (y - (a + 0.5 * torch.sin(2.5 * x + 3 * b))) ** 2 + (z - (c + 0.5 * torch.cos(2.0 * x + 3 * d))) ** 2
            vol += 45000.0 * torch.exp(-dist2 / (2 * 0.006 ** 2))
        vol += 50000.0 * torch.exp(-((x - 0.1) ** 2 + (y + 0.2) ** 2 + (z - 0.3) ** 2) / (2 * 0.05 ** 2))
    elif kind == "hipct":
        vol = torch.zeros(shape, device=device, dtype=torch.float32)
        ph = _params(6, 3, seed)
        for octave in range(6):
            k = 2.0 ** (octave + 2) * 3.14159265
            pz, py, px = [3.14159265 * (1 + v) for v in ph[octave]]
            vol = vol + (0.6 ** octave) * torch.sin(k * z + pz) * torch.sin(k * y + py) * torch.sin(k * x + px)
        contrast = 0.35 + 0.65 * (0.5 + 0.5 * torch.tanh(3 * (x + y * 0.5 - 0.3 * z)))
        vol = 30000.0 + 12000.0 * vol * contrast + _noise(shape, 400.0, seed, block_id, device)
    elif kind == "vessel":
        vol = 200.0 + _noise(shape, 50.0, seed, block_id, device)
        for a, b, c, d in _params(12, 4, seed):
            dist2 = (y - (a + 0.4 * torch.sin(2.5 * (1.5 + b) * x + 3 * z + b))) ** 2 + (z - (c * 0.6 + 0.3 * torch.cos(2.5 * (1.5 + d) * x + d))) ** 2
            vol += 30000.0 * torch.exp(-dist2 / (2 * 0.04 ** 2))
    else:
        raise KeyError(kind)
    return vol.clamp_(0, 65535).to(torch.int32).to(torch.int16).contiguous()  # uint16 bit pattern
