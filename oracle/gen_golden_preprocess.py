"""Golden vectors for `preprocess` (utils/misc.py:244-254): inputs and outputs of the UNMODIFIED reference function
(imported through oracle/refshim.py; it calls scipy.ndimage.binary_opening), with the oracle's two restatements
(brief_oracle.preprocess = the same scipy call, brief_oracle.preprocess_restated = the box opening written out)
checked bit for bit against it before anything is saved.  Run in the build container only:
    python oracle/gen_golden_preprocess.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import brief_oracle as O  # noqa: E402
import refshim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
ref = refshim.load_reference()


def noisy(shape, seed, dtype, floor, noise, blob_frac):
    """Dark noisy background (values around `floor`) with brighter structure: thresholding at ~floor gives a mask
    with isolated specks (removed by the opening) and solid regions (kept)."""
    rng = np.random.default_rng(seed)
    v = rng.normal(floor, noise, shape)
    grids = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    r2 = sum((g - c) ** 2 for g, c in zip(grids, rng.uniform(-0.5, 0.5, len(shape))))
    v += (r2 < blob_frac) * rng.uniform(4 * floor, 12 * floor)
    v[rng.random(shape) < 0.03] = 0
    tmax = np.iinfo(dtype).max
    return np.clip(v, 0, tmax).astype(dtype)[..., None]


cases = {
    # name: (data, level, close, clip)
    "shipped_noop": (noisy((9, 20, 37), 1, np.uint16, 300, 80, 0.3), 0, [2, 2, 2], [0, 65535]),
    "open222": (noisy((12, 33, 70), 2, np.uint16, 300, 80, 0.3), 300, [2, 2, 2], [0, 65535]),
    "open222_clip": (noisy((7, 24, 64), 3, np.uint16, 300, 80, 0.3), 310, [2, 2, 2], [150, 2000]),
    "open323": (noisy((10, 19, 45), 4, np.uint16, 500, 200, 0.2), 800, [3, 2, 3], [0, 65535]),
    "open144": (noisy((3, 40, 40), 5, np.uint16, 500, 200, 0.2), 800, [1, 4, 4], [0, 60000]),
    "open411": (noisy((16, 8, 33), 6, np.uint16, 500, 200, 0.2), 500, [4, 1, 1], [0, 65535]),
    "plain_threshold": (noisy((5, 17, 31), 7, np.uint16, 300, 80, 0.3), 290, False, [0, 65535]),
    "u8_open222": (noisy((8, 32, 48), 8, np.uint8, 20, 8, 0.3), 20, [2, 2, 2], [0, 255]),
    "u8_open232_clip": (noisy((6, 21, 50), 9, np.uint8, 20, 8, 0.3), 22, [2, 3, 2], [5, 200]),
    "thin_depth": (noisy((1, 30, 64), 10, np.uint16, 300, 80, 0.3), 320, [2, 2, 2], [0, 65535]),   # boxes never fit in z
    "image2d": (noisy((37, 53), 11, np.uint16, 300, 80, 0.3), 300, [2, 2, 2], [0, 65535]),         # [H,W,1]: close[:2]
    "all_below": (noisy((4, 9, 40), 12, np.uint16, 300, 80, 0.3), 65535, [2, 2, 2], [0, 65535]),
}

out = {}
for name, (data, level, close, clip) in cases.items():
    want = ref.misc.preprocess(data.copy(), level, close, clip)
    a = O.preprocess(data.copy(), level, close, clip)
    b = O.preprocess_restated(data, level, close, clip)
    assert want.dtype == data.dtype == a.dtype == b.dtype and want.shape == data.shape
    assert want.tobytes() == a.tobytes(), name
    assert want.tobytes() == b.tobytes(), name
    changed = int((want != data).sum())
    print(f"{name:18s} shape {data.shape} level {level} close {close} clip {clip}: {changed} voxels changed")
    out[name + "/in"] = data
    out[name + "/out"] = want
    out[name + "/level"] = np.int64(level)
    out[name + "/close"] = np.array([0, 0, 0] if close is False else close, np.int64)
    out[name + "/clip"] = np.array(clip, np.int64)
np.savez_compressed(os.path.join(GOLD, "preprocess.npz"), **out)
print("wrote", os.path.join(GOLD, "preprocess.npz"), os.path.getsize(os.path.join(GOLD, "preprocess.npz")), "bytes")
