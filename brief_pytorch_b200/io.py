"""Normalisation side-info — same signatures as the reference's utils/io.py:65-80, 111-147
(`minmaxany_lo_hi` family, the only one the shipped configs use).

In the fused path these run ON CHIP (csrc/brief_common.cuh: brief_normalize / brief_denorm); the host
versions below exist for the side-info (`sideinfos.yaml`) and for callers that hold numpy data.
"""
from typing import Tuple

import numpy as np
import torch

_INT_DTYPES = {"uint8": np.uint8, "uint16": np.uint16, "int16": np.int16, "float32": np.float32,
               "float64": np.float64}


def _scale(name: str) -> Tuple[float, float]:
    if "minmaxany" not in name:
        raise NotImplementedError(f"normalisation '{name}' is outside the SIREN hot path")
    _, lo, hi = name.split("_")
    return float(lo), float(hi)


def normalize_data(data: np.ndarray, name: str, min=None, max=None):
    """fp32: ((x - min) / (max - min)) * (hi - lo) + lo ; returns (tensor, sideinfos)."""
    lo, hi = _scale(name)
    dtype = data.dtype.name
    x = data.astype(np.float32)
    vmin = float(x.min()) if min is None else min
    vmax = float(x.max()) if max is None else max
    x = (x - vmin) / (vmax - vmin)
    x *= (hi - lo)
    x += lo
    t = torch.tensor(x, dtype=torch.float)
    info = {"dtype": dtype, "min": vmin, "max": vmax,
            "normalized_min": t.min().item(), "normalized_max": t.max().item()}
    return t, info


def invnormalize_data(data: torch.Tensor, sideinfos: dict, name: str) -> np.ndarray:
    """clip((y - lo)/(hi - lo), 0, 1) * (max - min) + min, then a TRUNCATING cast to the original dtype."""
    lo, hi = _scale(name)
    y = data.detach().to("cpu", torch.float32).clone()
    y -= lo
    y /= (hi - lo)
    y = torch.clip(y, 0, 1)
    y = y * (sideinfos["max"] - sideinfos["min"]) + sideinfos["min"]
    return np.array(y, dtype=_INT_DTYPES[sideinfos["dtype"]])


def normalized_threshold(weight_thres: float, name: str, vmin: float, vmax: float) -> float:
    """main.py:380-383: the raw weight threshold pushed through the block's normalisation."""
    t, _ = normalize_data(np.array(weight_thres), name, min=vmin, max=vmax)
    return float(t)


def get_type_max(data: np.ndarray) -> int:
    table = {"uint8": 255, "uint12": 4098, "uint16": 65535, "float32": 65535, "float64": 65535, "int16": 65535}
    if data.dtype.name not in table:
        raise NotImplementedError(data.dtype.name)
    return table[data.dtype.name]
