"""`preprocess` (utils/misc.py:244-254): the oracle's two restatements against golden vectors made by the
reference's own function (oracle/gen_golden_preprocess.py), and — on the GPU — brief_preprocess through the C-ABI,
bit for bit: golden cases, random shapes / structures / dtypes (vector and scalar row paths), and the properties an
opening has at any size (idempotent, anti-extensive, only mask voxels change)."""
import numpy as np
import pytest
import torch

import brief_oracle as O
from conftest import load_gold

CASES = ["shipped_noop", "open222", "open222_clip", "open323", "open144", "open411", "plain_threshold", "u8_open222",
         "u8_open232_clip", "thin_depth", "image2d", "all_below"]


def case(g, name):
    close = [int(c) for c in g[name + "/close"]]
    return g[name + "/in"], int(g[name + "/level"]), (False if close == [0, 0, 0] else close), [int(c) for c in g[name + "/clip"]]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = load_gold("preprocess")
    data, level, close, clip = case(g, name)
    want = g[name + "/out"]
    assert O.preprocess(data.copy(), level, close, clip).tobytes() == want.tobytes()       # scipy call, like the reference
    assert O.preprocess_restated(data, level, close, clip).tobytes() == want.tobytes()     # box opening written out


def test_box_opening_equals_scipy_for_every_structure_size():
    from scipy import ndimage
    rng = np.random.default_rng(3)
    for trial in range(40):
        shape = tuple(int(x) for x in rng.integers(1, 12, 3))
        size = tuple(int(x) for x in rng.integers(1, 5, 3))
        m = rng.random(shape) < rng.uniform(0.4, 0.95)
        want = ndimage.binary_opening(m, structure=np.ones(size), iterations=1)
        assert np.array_equal(O.box_opening(m, size), want), (shape, size)


def test_config_errors_follow_reference():
    from brief_pytorch_b200 import misc
    data = np.zeros((2, 3, 4, 1), np.uint16)
    with pytest.raises(AssertionError):  # range_limit, utils/tool.py:26-30
        misc.preprocess(data, 0, [2, 2, 2], [0, 70000])
    with pytest.raises(AssertionError):
        misc.preprocess(data, 0, [2, 2, 2], [10, 5])
    assert misc.preprocess_is_identity(np.uint16, 0, [0, 65535])      # every shipped yaml
    assert not misc.preprocess_is_identity(np.uint16, 1, [0, 65535])
    assert not misc.preprocess_is_identity(np.uint16, 0, [1, 65535])
    assert not misc.preprocess_is_identity(np.uint8, 0, [0, 200])


# ---- GPU: the CUDA kernels through the C-ABI ----------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(name):
    from brief_pytorch_b200 import misc
    g = load_gold("preprocess")
    data, level, close, clip = case(g, name)
    keep = data.copy()
    got = misc.preprocess(data, level, close, clip)
    assert got.dtype == data.dtype and got.shape == data.shape
    assert got.tobytes() == g[name + "/out"].tobytes()
    assert np.array_equal(data, keep)


@pytest.mark.gpu
def test_kernel_matches_oracle_on_random_shapes():
    """Rows that take the 16-byte vector path (W % 8 == 0 / W % 16 == 0) and rows that take the scalar path, every
    structure size 1..4 per axis, both dtypes, with and without a real clip."""
    from brief_pytorch_b200 import misc
    rng = np.random.default_rng(11)
    for trial in range(48):
        dtype = np.uint16 if trial % 3 else np.uint8
        tmax = np.iinfo(dtype).max
        w = int(rng.choice([8, 16, 24, 31, 32, 33, 64, 72, 95, 96, 128, 130]))
        shape = (int(rng.integers(1, 9)), int(rng.integers(1, 20)), w)
        close = False if trial % 8 == 7 else [int(x) for x in rng.integers(1, 5, 3)]
        floor = 40 if dtype == np.uint8 else 900
        data = np.clip(rng.normal(floor, floor / 4, shape), 0, tmax).astype(dtype)[..., None]
        level = int(floor * rng.uniform(0.9, 1.6))
        clip = [0, tmax] if trial % 2 else [int(floor * 0.5), int(floor * 1.4)]
        want = O.preprocess_restated(data, level, close, clip)
        got = misc.preprocess(data, level, close, clip)
        assert got.tobytes() == want.tobytes(), (trial, shape, dtype, close, level, clip)


@pytest.mark.gpu
def test_kernel_properties_at_block_size():
    """A 64 x 256 x 256 block (the vessel configuration's): the result is idempotent, never raises a value, changes
    only voxels at or below the level, and agrees with the oracle on a sub-volume cut far from the faces."""
    from brief_pytorch_b200.group import preprocess_
    from brief_pytorch_b200 import synth
    vol = synth.vessel((64, 256, 256), seed=3)[..., 0]
    level = int(np.quantile(vol, 0.6))
    t = torch.from_numpy(vol.view(np.int16)).cuda()
    preprocess_(t, level, [2, 2, 2], [0, 65535], "uint16")
    once = t.cpu().numpy().view(np.uint16)
    preprocess_(t, level, [2, 2, 2], [0, 65535], "uint16")
    twice = t.cpu().numpy().view(np.uint16)
    changed = once != vol
    assert changed.sum() > 1000 and (once <= vol).all() and (vol[changed] <= level).all() and (once[changed] == 0).all()
    # idempotent: zeroed voxels stay in the mask, so the second opening contains the first and changes nothing else
    assert np.array_equal(once, twice)
    want = O.preprocess_restated(vol[..., None], level, [2, 2, 2], [0, 65535])[..., 0]
    assert np.array_equal(once, want)


@pytest.mark.gpu
def test_error_paths():
    import ctypes as C
    from brief_pytorch_b200 import _cabi
    from brief_pytorch_b200.group import preprocess_
    t = torch.zeros((4, 8, 16), dtype=torch.int16, device="cuda")
    with pytest.raises(_cabi.BriefError):
        preprocess_(t, 10, [5, 2, 2], [0, 65535], "uint16")      # structure side outside 1..4
    with pytest.raises(_cabi.BriefError):
        preprocess_(t, 10, [2, 2, 2], [0, 65536], "uint16")      # Improper range setting!
    with pytest.raises(_cabi.BriefError):
        preprocess_(t, 10, [2, 2, 2], [9, 3], "uint16")
    lib = _cabi.load()
    assert lib.brief_preprocess(None, 1, 4, 8, 16, 0.0, None, 0.0, 65535.0, None, 0, None) < 0
    assert lib.brief_preprocess_scratch_bytes(4, 8, 16) == 2 * 4 * 8 * 1 * 4


@pytest.mark.gpu
def test_framework_applies_pre_and_postprocess_per_block(tmp_path):
    """Compress.preprocess runs on every block before min / max are taken (main.py:336-342) and
    Decompress.postprocess on every decoded block (main.py:295), both per block, both on the device."""
    import os
    import yaml
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    from test_framework import opt
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    vol = synth.vessel((16, 48, 48), seed=7)
    level = int(np.quantile(vol, 0.5))
    o["Compress"]["preprocess"] = {"denoise": {"level": level, "close": [2, 2, 2]}, "clip": [0, 30000]}
    o["Decompress"]["postprocess"] = {"denoise": {"level": level, "close": [2, 2, 2]}, "clip": [0, 30000]}
    cf = NFGR(o, 0, "f16")
    cdir = str(tmp_path / "compressed")
    blocks, _ = cf.compress_divide(vol, cdir, max_steps=40)
    for b in blocks:  # sideinfos min / max are those of the preprocessed block
        pre = O.preprocess_restated(b.data, level, [2, 2, 2], [0, 30000])
        assert (b.sideinfos["min"], b.sideinfos["max"]) == (float(pre.min()), float(pre.max()))
    ours = cf.decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"),
                                os.path.join(cdir, "sideinfos"))
    # decode the same directory without the postprocess, apply the oracle's postprocess per block, merge
    o2 = opt()
    o2["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    plain = NFGR(o2, 0, "f16").decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"),
                                                  os.path.join(cdir, "sideinfos"))
    want = np.zeros_like(plain)
    for b in blocks:
        sl = (slice(b.d[0], b.d[1] + 1), slice(b.h[0], b.h[1] + 1), slice(b.w[0], b.w[1] + 1))
        want[sl] = O.preprocess_restated(plain[sl], level, [2, 2, 2], [0, 30000])
    assert np.array_equal(ours, want)
    assert (ours != plain).sum() > 0
