"""Network constructor — drop-in for the SIREN part of the reference's utils/Networks.py
(`SIREN` :246-314, `Sine` :227-234, `init_phi` :800-802, `get_nnmodule_param_count` :13-17).

The module keeps the reference's parameter tree (`net[l][0].weight` [out,in], `net[l][0].bias`), so
`save_model` / `load_model` / `state_dict` are interchangeable, and it consumes torch's CPU generator in
the reference's order, so the initial weights are bit-identical under the same seed.  What differs is
`forward`: on a CUDA module it runs the fused sm_100a kernels of libbrief_b200 (no per-layer ATen ops);
on a CPU module it raises — this package has no CPU compute path.
"""
from __future__ import annotations

import copy
import math
from typing import Mapping

import numpy as np
import torch
from torch import nn

HIDDEN_W0 = 30  # every Sine() after the first layer (reference default, never configured)


def get_nnmodule_param_count(module: nn.Module) -> int:
    return sum(int(np.prod(t.shape)) for t in module.state_dict().values())


class Sine(nn.Module):
    """y = sin(w0 * x).  Kept as a (parameter-free) module so the module tree matches the reference."""

    def __init__(self, w0: float = HIDDEN_W0):
        super().__init__()
        self.w0 = w0

    def forward(self, x):  # only reached if a caller walks the Sequential by hand on CUDA tensors
        return torch.sin(self.w0 * x)


def _uniform_(weight: torch.Tensor, bound: float) -> None:
    with torch.no_grad():
        weight.uniform_(-bound, bound)


class SIREN(nn.Module):
    """SIREN(coords_channel, data_channel, features, layers, w0): `layers` Linear maps, sine after all but
    the last; first-layer frequency w0, hidden frequency 30."""

    def __init__(self, coords_channel=3, data_channel=1, features=256, layers=5, w0=30, res=False,
                 output_act=False, **kwargs):
        super().__init__()
        if res:
            raise NotImplementedError("res=True (HalfResidual blocks) is not part of the fused SIREN path")
        if output_act:
            raise NotImplementedError("output_act=True is not part of the fused SIREN path")
        if layers < 2:
            raise ValueError("SIREN needs at least 2 layers")
        self.coords_channel, self.data_channel = int(coords_channel), int(data_channel)
        self.features, self.layers, self.w0 = int(features), int(layers), float(w0)
        widths = [self.coords_channel] + [self.features] * (self.layers - 1) + [self.data_channel]
        blocks = []
        for l in range(self.layers):  # nn.Linear's own init draws (weight, bias) per layer, in layer order
            lin = nn.Linear(widths[l], widths[l + 1])
            blocks.append(nn.Sequential(lin, Sine(w0 if l == 0 else HIDDEN_W0)) if l < self.layers - 1
                          else nn.Sequential(lin))
        self.net = nn.Sequential(*blocks)
        for blk in self.net:  # then every weight is redrawn U(+-sqrt(6/fan_in)/30), in layer order ...
            _uniform_(blk[0].weight, np.sqrt(6 / blk[0].weight.size(-1)) / 30)
        _uniform_(self.net[0][0].weight, 1 / self.net[0][0].weight.size(-1))  # ... and layer 0 once more
        self.precision = kwargs.get("precision", "auto")
        self._group = None
        self._group_key = None

    # -- fused evaluation ------------------------------------------------------------------------------
    def _param_key(self):
        ps = list(self.parameters())
        return (ps[0].device, self.precision, tuple((p.data_ptr(), p._version) for p in ps))

    def fused_group(self, dims=(1, 1, 1)):
        """Single-network SirenGroup mirroring this module's current parameters (rebuilt lazily)."""
        from .group import NetSpec, SirenGroup
        dims = tuple(int(d) for d in dims)
        key = (self._param_key(), dims)
        if self._group is None or self._group_key != key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("brief_pytorch_b200.SIREN.forward needs the module on a CUDA device: "
                                   "this package has no CPU compute path")
            if self._group is not None:
                self._group.close()
            spec = NetSpec(self.features, self.layers, self.w0, dims if self.coords_channel == 3 else dims[-2:],
                           self.coords_channel, self.data_channel, float(HIDDEN_W0))
            self._group = SirenGroup([spec], dev, self.precision)
            self._group.load_module(0, self)
            self._group_key = key
        return self._group

    def forward(self, coords: torch.Tensor) -> torch.Tensor:
        if coords.dtype not in (torch.float32, torch.float16):
            raise TypeError(f"coords must be float32 (or float16 after .half()), got {coords.dtype}")
        if not coords.is_cuda:
            raise RuntimeError("brief_pytorch_b200 has no CPU compute path: move the module and coords to CUDA")
        dims = self._group_key[1] if self._group_key is not None else (1, 1, 1)
        with torch.no_grad():
            # module.half() + half coordinates (main.py:388-391, utils/misc.py:83-84): the fused kernels evaluate the given
            # (fp16-representable) values in their own arithmetic; the result goes back in the caller's dtype
            return self.fused_group(dims).forward(0, coords.float()).to(coords.dtype)

    def __deepcopy__(self, memo):
        clone = SIREN(self.coords_channel, self.data_channel, self.features, self.layers, self.w0,
                      precision=self.precision)
        clone.load_state_dict(copy.deepcopy(self.state_dict(), memo))
        return clone.to(next(self.parameters()).device)

    # -- byte budget <-> width (closed forms of the reference's static methods) ---------------------------
    @staticmethod
    def calc_param_count(coords_channel, data_channel, features, layers, res=False, **kwargs) -> int:
        hidden = (layers - 2) * (features * features + features)
        if res:
            hidden *= 2
        return int(coords_channel * features + features + hidden + features * data_channel + data_channel)

    @staticmethod
    def calc_features(param_count, coords_channel, data_channel, layers, res=False, **kwargs) -> int:
        # P = a f^2 + b f + data_channel  ->  positive root, rounded to the nearest integer
        a = (layers - 2) * (2 if res else 1)
        b = coords_channel + 1 + (2 * layers - 4 if res else layers - 2) + data_channel
        c = data_channel - param_count
        if a == 0:
            return round(-c / b)
        return round((math.sqrt(b * b - 4 * a * c) - b) / (2 * a))


ALLPHI = {"SIREN": SIREN}
ALL_CALC_PHI_FEATURES = {"SIREN": SIREN.calc_features}
ALL_CALC_PHI_PARAM_COUNT = {"SIREN": SIREN.calc_param_count}
ALL_CHECK_PARAM_COUNT = {}


def init_phi(kwargs: Mapping) -> SIREN:
    """init_phi({'name': 'SIREN', ...}); unknown names raise KeyError like the reference's dict lookup."""
    kw = copy.deepcopy(dict(kwargs))
    return ALLPHI[kw.pop("name")](**kw)
