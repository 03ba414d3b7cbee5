"""Golden vectors for the deblocking filter: inputs, block order and outputs of the UNMODIFIED reference deblock.cpp
(built by oracle/build_ref.sh).  Run in the build container:  python oracle/gen_golden_deblock.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import deblock_oracle as D  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
assert D.available(), "run oracle/build_ref.sh first"


def blocky(shape, grid, seed, step=120, base=20000, noise=30):
    """Smooth volume + a per-block offset (what independent per-block fits leave behind) + a little noise."""
    rng = np.random.default_rng(seed)
    d, h, w = shape
    zz, yy, xx = np.meshgrid(np.arange(d), np.arange(h), np.arange(w), indexing="ij")
    v = base + 3000 * np.sin(zz / 5.0 + yy / 9.0) * np.cos(xx / 7.0)
    gd, gh, gw = grid
    names = []
    for iz in range(gd):
        for iy in range(gh):
            for ix in range(gw):
                z1, z2 = iz * d // gd, (iz + 1) * d // gd - 1
                y1, y2 = iy * h // gh, (iy + 1) * h // gh - 1
                x1, x2 = ix * w // gw, (ix + 1) * w // gw - 1
                v[z1:z2 + 1, y1:y2 + 1, x1:x2 + 1] += rng.integers(-step, step + 1)
                names.append(D.block_name(z1, z2, y1, y2, x1, x2))
    v += rng.integers(-noise, noise + 1, size=shape)
    return np.clip(v, 0, 65535).astype(np.uint16), names


cases = {
    "a": blocky((6, 40, 48), (1, 2, 2), 1),
    "b": blocky((9, 48, 60), (3, 2, 3), 2, step=200),
    "c": blocky((4, 33, 35), (2, 3, 5), 3, step=60, base=300, noise=250),   # dark: uint16 wrap-around of p +/- delta
    "d": blocky((5, 20, 64), (1, 1, 8), 4, step=400, noise=5),
}
out = {}
for tag, (vol, names) in cases.items():
    res, order = D.run_reference(vol, names)
    out[f"{tag}_in"] = vol
    out[f"{tag}_out"] = res
    out[f"{tag}_order"] = np.array(order)
    print(tag, vol.shape, len(names), "blocks, changed voxels:", int((res != vol).sum()))
np.savez_compressed(os.path.join(GOLD, "deblock.npz"), **out)
