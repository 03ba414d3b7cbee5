"""The NFGR framework layer (brief_pytorch_b200/CompressFramework.py) against the reference's semantics
(main.py:199-246, 484-532, 589-607 restated in oracle/brief_oracle.py): byte budget -> width, partition, budget
allocation, the compressed/ directory layout, and — on the GPU — fit + decode of a divided volume whose written
directory the ORACLE decodes to the same voxels."""
import copy
import os

import numpy as np
import pytest
import torch
import yaml

import brief_oracle as O

VESSEL_YAML = """
Name: NFGR
Compress:
  divide: {divide_type: adaptotal_-1_-1_-1_4, param_alloc: by_size, param_size_thres: 26, exception: none}
  half: false
  sampler: {name: randomcube, cube_count: 1, cube_len: [10000000, 10000000, 10000000], sample_size: 100000}
  coords_mode: -1,1
  preprocess: {denoise: {level: 0, close: [2, 2, 2]}, clip: [0, 65535]}
  param: {init_net_path: none, filesize_ratio: 128, given_size: 0}
  loss: {name: datal2, beta: 0.01, weight: [value_65535_65535_1], weight_thres: 65535}
  max_steps: 80000
  checkpoints: every_2000
  lr_phi: 0.001
  optimizer_name_phi: Adamax
  lr_scheduler_phi: {name: MultiStepLR, milestones: [50000, 60000, 70000], gamma: 0.2}
Decompress:
  sample_size: 10000
  postprocess: {denoise: {level: 0, close: [2, 2, 2]}, clip: [0, 65535]}
Module:
  phi: {coords_channel: 3, data_channel: 1, layers: 7, name: SIREN, w0: 10, output_act: false, res: false}
Normalize: {name: minmaxany_0_100}
"""


def opt():
    return yaml.safe_load(VESSEL_YAML)


def test_budget_width_and_partition_match_reference_sizing():
    from brief_pytorch_b200.CompressFramework import NFGR
    cf = NFGR(opt())
    # vessel as shipped -> 4 blocks 64x256x256, f = 56, 16241 parameters
    vol = np.zeros((64, 512, 512, 1), np.uint16)
    ps = cf.parse_param_size(vol.nbytes)
    assert ps == vol.nbytes / 128
    blocks = cf.divide(vol, ps)
    assert [b.name for b in blocks] == ["d_0_63-h_0_255-w_0_255", "d_0_63-h_0_255-w_256_511",
                                        "d_0_63-h_256_511-w_0_255", "d_0_63-h_256_511-w_256_511"]
    for b in blocks:
        f, nbytes = cf.estimate_module_size(b.param_size)
        assert (f, nbytes) == (56, 16241 * 4.0)
        assert f == O.calc_features(b.param_size / 4.0, 3, 1, 7)
    # Nb = 64 -> 64^3 blocks, f = 13
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "adaptotal_-1_-1_-1_64"
    cf = NFGR(o)
    blocks = cf.divide(vol, ps)
    assert len(blocks) == 64 and blocks[0].shape == (64, 64, 64)
    assert cf.estimate_module_size(blocks[0].param_size)[0] == 13


def test_config_errors_follow_reference():
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["param"]["given_size"] = 100
    with pytest.raises(ValueError):  # main.py:200-201
        NFGR(o).parse_param_size(1000)
    o = opt()
    o["Compress"]["loss"]["name"] = "nope"
    with pytest.raises(NotImplementedError):  # main.py:197
        NFGR(o)
    o = opt()
    o["Module"]["phi"]["name"] = "NeRF"
    with pytest.raises(KeyError):  # utils/Networks.py:800-802
        NFGR(o)
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "adaptive_4_2_0.1_0.1_16"
    with pytest.raises(NotImplementedError):
        NFGR(o).divide(np.zeros((8, 8, 8, 1), np.uint16), 100.0)


def test_by_var_allocation_matches_oracle():
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"].update(divide_type="total_2_2_2", param_alloc="by_var")
    vol = synth.hipct((32, 32, 32), seed=5)
    blocks = NFGR(o).divide(vol, 8 * 4000.0)
    ref = O.alloc_param(O.divide_data(vol, "total_2_2_2"), 8 * 4000.0, "by_var", 26)
    assert [b.name for b in blocks] == [c["name"] for c in ref]
    np.testing.assert_allclose([b.param_size for b in blocks], [c["param_size"] for c in ref], rtol=0, atol=0)
    assert len({round(b.param_size) for b in blocks}) > 1  # non-uniform budgets


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_divided_fit_writes_a_directory_the_oracle_decodes(tmp_path, prec):
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    vol = synth.vessel((16, 48, 48), seed=7)
    cf = NFGR(o, 0, prec)
    cdir = str(tmp_path / "compressed")
    blocks, mine = cf.compress_divide(vol, cdir, max_steps=60)
    assert mine == list(range(4)) and all(np.isfinite(b.loss) for b in blocks)
    # directory layout of main.py:589-607
    assert sorted(os.listdir(cdir)) == ["module", "sideinfos", "sideinfos.yaml"]
    for b in blocks:
        files = sorted(os.listdir(os.path.join(cdir, "module", b.name, "module")))
        assert len(files) == 14 and files[0].startswith("bias-0-")
    ours = cf.decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"),
                                os.path.join(cdir, "sideinfos"))
    assert ours.shape == vol.shape and ours.dtype == vol.dtype
    # the oracle (reference semantics, torch CPU fp32) decodes the same directory
    chunks = []
    for b in blocks:
        with open(os.path.join(cdir, "sideinfos", b.name, "sideinfos.yaml")) as fh:
            side = yaml.safe_load(fh)
        m = O.init_phi(dict(o["Module"]["phi"], features=side["phi_features"]))
        O.load_model(m, os.path.join(cdir, "module", b.name, "module"))
        chunks.append({"data": O.decompress_block(m, side, "minmaxany_0_100"), "name": b.name, "d": b.d, "h": b.h, "w": b.w})
    ref = O.merge_divided_data(chunks, list(vol.shape))
    span = float(vol.max()) - float(vol.min())
    tol = 1e-4 if prec == "fp32" else 1e-2
    d = np.abs(ours.astype(np.int64) - ref.astype(np.int64)).max()
    assert d <= np.ceil(3 * tol * span) + 1, d
    # and the fit did something: PSNR of ours vs the volume within 0.1 dB of the oracle's decode of the same weights
    p_ours, p_ref = O.cal_psnr(vol, ours, 65535), O.cal_psnr(vol, ref, 65535)
    assert abs(p_ours - p_ref) < 0.1, (p_ours, p_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("ratio,lo,hi", [(0.4, 30, 80), (0.13, 50, 140)])
def test_by_var_group_mixes_fused_and_wide_kernels(tmp_path, ratio, lo, hi):
    """hipct.yaml's allocation (budgets proportional to block variance) gives every block its own width: 25 .. 88 in the
    first case (two buckets of the fused tensor-core fit kernel and two of the wide one in ONE group), up to ~150 in the
    second (the layer-wise kernels of brief_tc_lw.cu join).  The written directory is decoded by the oracle to the same
    voxels (f16 tolerance) and the fit lowered every block's loss."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"].update(divide_type="total_2_2_2", param_alloc="by_var")
    o["Compress"]["param"]["filesize_ratio"] = ratio
    o["Compress"]["checkpoints"] = "none"
    vol = synth.hipct((32, 64, 64), seed=5)
    cf = NFGR(o, 0, "auto")
    cdir = str(tmp_path / "compressed")
    first, _ = NFGR(copy.deepcopy(o), 0, "auto").compress_divide(vol, None, max_steps=1)
    blocks, _ = cf.compress_divide(vol, cdir, max_steps=60)
    widths = sorted({b.features for b in blocks})
    assert widths[0] <= lo and widths[-1] >= hi, widths
    for b0, b in zip(first, blocks):
        assert np.isfinite(b.loss) and b.loss < b0.loss
    ours = cf.decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"),
                                os.path.join(cdir, "sideinfos"))
    chunks = []
    for b in blocks:
        with open(os.path.join(cdir, "sideinfos", b.name, "sideinfos.yaml")) as fh:
            side = yaml.safe_load(fh)
        m = O.init_phi(dict(o["Module"]["phi"], features=side["phi_features"]))
        O.load_model(m, os.path.join(cdir, "module", b.name, "module"))
        chunks.append({"data": O.decompress_block(m, side, "minmaxany_0_100"), "name": b.name, "d": b.d, "h": b.h, "w": b.w})
    ref = O.merge_divided_data(chunks, list(vol.shape))
    span = float(vol.max()) - float(vol.min())
    d = np.abs(ours.astype(np.int64) - ref.astype(np.int64)).max()
    assert d <= np.ceil(3 * 1e-2 * span) + 1, d


@pytest.mark.gpu
@pytest.mark.parametrize("shape,steps,ratio", [((16, 48, 48), 20, 16), ((96, 192, 192), 6, 2048), ((16, 48, 48), 8, 0.15)])
def test_block_ownership_does_not_change_a_blocks_result(shape, steps, ratio):
    """with per-network slicing (reproducible=True) a block's fitted parameters are bit-identical
    whichever rank owns it / whatever shares the GPU.  The second case has random-point blocks of 96x96x96 voxels
    (batch 100000 -> several slices per network, on-device sampler keyed by the block's global index); the third has
    f = 77 networks, i.e. the wide tensor-core fit kernel."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    from brief_pytorch_b200.group import pack_module_params
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = ratio
    o["Compress"]["checkpoints"] = "none"
    vol = synth.vessel(shape, seed=7)
    all_blocks, _ = NFGR(o, 0, "f16", reproducible=True).compress_divide(vol, None, max_steps=steps)
    got = {}
    for rank in range(2):
        blocks, mine = NFGR(copy.deepcopy(o), 0, "f16", reproducible=True).compress_divide(vol, None, max_steps=steps,
                                                                                       rank=rank, world=2)
        for i in mine:
            got[i] = pack_module_params(blocks[i].module)
    assert sorted(got) == [0, 1, 2, 3]
    for i in range(4):
        np.testing.assert_array_equal(got[i], pack_module_params(all_blocks[i].module))
        assert np.isfinite(got[i]).all()


@pytest.mark.gpu
def test_warm_start_from_a_compressed_directory(tmp_path):
    """Compress.param.init_net_path (main.py:345-354): a second run that starts from the first run's modules begins
    at the first run's final loss instead of the initial one."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    vol = synth.vessel((16, 48, 48), seed=7)
    cdir = str(tmp_path / "compressed")
    cold, _ = NFGR(o, 0, "f16").compress_divide(vol, cdir, max_steps=150)
    first, _ = NFGR(o, 0, "f16").compress_divide(vol, None, max_steps=1)
    o["Compress"]["param"]["init_net_path"] = os.path.join(cdir, "module")
    warm, _ = NFGR(o, 0, "f16").compress_divide(vol, None, max_steps=1)
    for c, f, w in zip(cold, first, warm):
        assert w.loss < f.loss - 1.0                 # clearly below a cold first step ...
        assert abs(w.loss - c.loss) < 0.01 * c.loss  # ... and where the first run stopped (one Adamax step apart)


def test_sampler_rules_follow_the_reference():
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    cf = NFGR(o, 0, "f16")
    assert cf._sampler_name(64 ** 3, (64, 64, 64)) == "randomcube"
    assert cf._sampler_name(96 ** 3, (96, 96, 96)) == "randompoint"       # main.py:332-334
    o["Compress"]["sampler"]["cube_len"] = [8, 8, 8]                      # min(block, cube) = 512 <= 80^3: stays a cube sampler
    assert NFGR(o, 0, "f16")._sampler_name(96 ** 3, (96, 96, 96)) == "randomcube"   # (sliding cubes: tests/test_cubes.py)
    o["Compress"]["sampler"]["name"] = "gridsampler"
    with pytest.raises(NotImplementedError):                             # main.py:371
        NFGR(o, 0, "f16")._sampler_name(64 ** 3, (64, 64, 64))


def test_exception_merge_and_volume_reader(tmp_path):
    """OmegaConf.merge semantics for Compress.divide.exception (main.py:568-569) and the .npy reader of NFGR.compress."""
    from brief_pytorch_b200.CompressFramework import _merge, read_volume
    base = {"Compress": {"lr_phi": 1e-3, "sampler": {"name": "randomcube", "sample_size": 5}}, "Module": {"phi": {"layers": 7}}}
    over = {"Compress": {"sampler": {"sample_size": 9}}, "Module": {"phi": {"layers": 5, "w0": 20}}}
    got = _merge(base, over)
    assert got == {"Compress": {"lr_phi": 1e-3, "sampler": {"name": "randomcube", "sample_size": 9}},
                   "Module": {"phi": {"layers": 5, "w0": 20}}}
    assert base["Module"]["phi"] == {"layers": 7}  # inputs untouched
    vol = (np.arange(2 * 3 * 4) * 7).astype(np.uint16).reshape(2, 3, 4)
    np.save(tmp_path / "v.npy", vol)
    r = read_volume(str(tmp_path / "v.npy"))
    assert r.shape == (2, 3, 4, 1) and r.dtype == np.uint16 and (r[..., 0] == vol).all()
    np.save(tmp_path / "w.npy", vol[..., None])
    assert read_volume(str(tmp_path / "w.npy")).shape == (2, 3, 4, 1)
