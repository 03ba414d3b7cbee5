"""Deblocking post-filter: the oracle's numpy restatement against golden vectors made by the reference's
own deblock.cpp (compiled unmodified by oracle/build_ref.sh with a stand-in tiffio.h), the host-side seam planner of the
C-ABI, and — on the GPU — the CUDA kernel, bit for bit."""
import os

import numpy as np
import pytest
import torch

import deblock_oracle as D
from conftest import load_gold

CASES = "abcd"


def test_restatement_matches_the_compiled_reference_golden():
    g = load_gold("deblock")
    for t in CASES:
        out = D.deblock_restated(g[t + "_in"], list(g[t + "_order"]))
        assert out.dtype == np.uint16 and out.tobytes() == g[t + "_out"].tobytes(), t
        assert (g[t + "_in"] != g[t + "_out"]).sum() > 1000  # the filter did something


@pytest.mark.skipif(not D.available(), reason="oracle/_ref/deblock_ref not built (needs the reference sources)")
def test_restatement_matches_the_compiled_reference_live():
    rng = np.random.default_rng(31)
    vol = (1500 + 400 * rng.random((3, 26, 31))).astype(np.uint16)
    names = [D.block_name(0, 2, 0, 12, 0, 14), D.block_name(0, 2, 0, 12, 15, 30), D.block_name(0, 2, 13, 25, 0, 30)]
    ref, order = D.run_reference(vol, names)
    assert D.deblock_restated(vol, order).tobytes() == ref.tobytes()


def test_seam_planner_of_the_library_matches_the_oracle():
    from brief_pytorch_b200.deblock import parse_chunk_name, seam_masks
    g = load_gold("deblock")
    for t in CASES:
        names = list(g[t + "_order"])
        assert seam_masks(names) == D.plan_lines(names)
    # sticky duplicate flags (deblock.cpp:244-276): a repeated block raises all four flags for every later block
    names = [D.block_name(0, 3, 0, 9, 0, 9), D.block_name(0, 3, 0, 9, 10, 19), D.block_name(0, 3, 0, 9, 0, 9),
             D.block_name(0, 3, 10, 19, 0, 9)]
    assert seam_masks(names) == D.plan_lines(names) == [15, 15, 0, 0]
    assert parse_chunk_name("d_0_63-h_256_511-w_0_255") == (0, 63, 256, 511, 0, 255)
    assert seam_masks([]) == []


@pytest.mark.gpu
def test_kernel_is_bit_identical_to_the_reference():
    from brief_pytorch_b200.deblock import deblock_
    g = load_gold("deblock")
    for t in CASES:
        vol = torch.from_numpy(g[t + "_in"].view(np.int16)).cuda()
        deblock_(vol, list(g[t + "_order"]))
        assert vol.cpu().numpy().view(np.uint16).tobytes() == g[t + "_out"].tobytes(), t


@pytest.mark.gpu
def test_kernel_at_volume_size_against_the_restatement(tmp_path):
    """64 x 256 x 256 decoded volume, 64 blocks (vessel Nb = 64 geometry): kernel == restatement; only voxels within
    two of a seam change; other thresholds than the defaults."""
    from brief_pytorch_b200.deblock import deblock_volume
    rng = np.random.default_rng(5)
    shape, grid = (64, 256, 256), (1, 8, 8)
    zz, yy, xx = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    vol = 12000 + 2000 * np.sin(yy / 17.0) * np.cos(xx / 13.0 + zz / 9.0)
    mod = tmp_path / "module"
    for iy in range(8):
        for ix in range(8):
            vol[:, iy * 32:(iy + 1) * 32, ix * 32:(ix + 1) * 32] += rng.integers(-150, 151)
            os.makedirs(mod / D.block_name(0, 63, iy * 32, iy * 32 + 31, ix * 32, ix * 32 + 31))
    vol = np.clip(vol + rng.integers(-20, 21, size=shape), 0, 65535).astype(np.uint16)
    for kw in ({}, {"index_a": 40, "index_b": 600, "thres": 12500}):
        got = deblock_volume(vol[..., None], str(mod), **kw)[..., 0]
        want = D.deblock_restated(vol, os.listdir(mod), **kw)
        assert got.tobytes() == want.tobytes()
        changed = np.argwhere(got != vol)
        assert len(changed) > 10000
        near = (np.minimum(changed[:, 1] % 32, 31 - changed[:, 1] % 32) <= 2) | (np.minimum(changed[:, 2] % 32, 31 - changed[:, 2] % 32) <= 2)
        assert near.all()
