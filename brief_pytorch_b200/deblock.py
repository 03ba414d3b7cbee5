"""Deblocking post-filter — the reference's `deblock <step_dir>` (deblock.cpp, C++/libtiff, single-threaded) as one CUDA
launch on the decoded volume, bit-identical to the C++ filter (integer arithmetic, in-place sequential seam order).

    vol = deblock_volume(decoded_uint16_dhw, module_dir)          # module_dir = .../compressed/module (chunk dirs)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import check


def parse_chunk_name(name: str) -> Tuple[int, int, int, int, int, int]:
    """'d_z1_z2-h_y1_y2-w_x1_x2' (main.py:589-607, inclusive ends) -> (z1, z2, y1, y2, x1, x2)."""
    parts = [p.split("_") for p in name.split("-")]
    return tuple(int(v) for p in parts for v in p[1:3])


def seam_masks(names: Sequence[str]) -> List[int]:
    """Which of its four seams (bit 0..3 = left, right, down, up) every block contributes, in listing order
    (deblock.cpp:244-276, sticky duplicate flags).  Host-only: no CUDA device needed."""
    lib = _cabi.load()
    n = len(names)
    blocks = (C.c_int32 * (6 * max(n, 1)))(*[v for nm in names for v in parse_chunk_name(nm)])
    masks = (C.c_int32 * max(n, 1))()
    check(lib.brief_deblock(None, 0, 0, 0, blocks, n, 51, 2000, 65535, masks, 0, None))
    return list(masks)[:n]


def deblock_(vol: torch.Tensor, names: Sequence[str], index_a: int = 51, index_b: int = 2000, thres: int = 65535) -> torch.Tensor:
    """In place on a CUDA tensor [D,H,W] holding uint16 voxels (int16 bit patterns or torch.uint16).  `names`: the chunk
    directory names in the order the reference would list them (os.listdir order of compressed/module)."""
    assert vol.is_cuda and vol.is_contiguous() and vol.dim() == 3 and vol.element_size() == 2
    lib = _cabi.load()
    n = len(names)
    blocks = (C.c_int32 * (6 * max(n, 1)))(*[v for nm in names for v in parse_chunk_name(nm)])
    d, h, w = (int(x) for x in vol.shape)
    with torch.cuda.device(vol.device):
        check(lib.brief_deblock(C.c_void_p(vol.data_ptr()), d, h, w, blocks, n, index_a, index_b, thres, None,
                                vol.device.index or 0, C.c_void_p(torch.cuda.current_stream(vol.device).cuda_stream)))
    return vol


def deblock_volume(volume: np.ndarray, module_dir: str, device: int = 0, **kw) -> np.ndarray:
    """The reference's CLI flow on arrays: decoded uint16 [D,H,W] (or [D,H,W,1]) + compressed/module -> filtered copy."""
    v = np.ascontiguousarray(volume.reshape(volume.shape[:3]), dtype=np.uint16)
    t = torch.from_numpy(v.view(np.int16)).to(f"cuda:{device}")
    deblock_(t, os.listdir(module_dir), **kw)
    return t.cpu().numpy().view(np.uint16).reshape(volume.shape)
