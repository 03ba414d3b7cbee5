"""Round-2 additions on the GPU, through the C-ABI: the graphed host step (brief_fit_step_host), the device histogram
behind the quantile weight rule, and the framework semantics the advisor flagged (whole-volume preprocess before the
partition, constant blocks, chunks_numbers, per-block exception overrides, NFGR.compress)."""
import copy
import os

import numpy as np
import pytest
import torch
import yaml

import brief_oracle as O
from test_framework import opt

pytestmark = pytest.mark.gpu


def _small_group(prec="f16", n=2, f=24, L=5, dims=(24, 40, 40), batch=3000):
    from brief_pytorch_b200 import Networks, synth
    from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
    grp = SirenGroup([NetSpec(f, L, 10.0, dims) for _ in range(n)], 0, prec)
    keep = []
    for j in range(n):
        blk = synth.neuron(dims, seed=3 + j)
        torch.manual_seed(42 + j)
        grp.set_params(j, pack_module_params(Networks.init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
        raw = torch.from_numpy(np.ascontiguousarray(blk[..., 0]).view(np.int16)).cuda()
        keep.append(raw)
        grp.bind_volume(j, raw, float(blk.min()), float(blk.max()), 0.0, 100.0, rules=[(10001, 65535, 0.1)], tau=40.0,
                        np_dtype="uint16")
        grp.set_sampler(j, "randompoint", batch)
    return grp, keep


@pytest.mark.parametrize("prec", ["f16", "fp32"])
@pytest.mark.parametrize("host_indices", [True, False])
def test_graphed_host_step_equals_separate_launches(prec, host_indices):
    """brief_fit_step_host (inputs on the copy stream, then one CUDA-graph launch: fit, optimiser, d2h loss) must report
    exactly the losses and leave exactly the parameters of brief_fit_step + brief_opt_step on the same indices / sampler
    stream.  After a MultiStepLR milestone the two paths form the learning rate differently (brief_opt_step takes it as
    a float, the graphed step keeps torch's chained double product), so from there on the comparison is to 1e-5; the
    device-sampler variant is also compared bit for bit with brief_fit_run, which shares the schedule arithmetic."""
    n, batch, steps, ms = 2, 3000, 7, [3, 5]
    a, keep_a = _small_group(prec, n, batch=batch)
    b, keep_b = _small_group(prec, n, batch=batch)
    n_vox = 24 * 40 * 40
    gen = torch.Generator().manual_seed(1)
    idx = [torch.empty(n * batch, dtype=torch.int64).pin_memory() for _ in range(2)]
    loss = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    for s in range(steps):
        k = s & 1
        torch.cuda.synchronize()
        if host_indices:
            idx[k].copy_(torch.randint(0, n_vox, (n * batch,), generator=gen))
        a.fit_step_host(idx[k] if host_indices else None, loss[k], "Adamax", 1e-3, milestones=ms, gamma=0.2, seed=42)
        want = b.fit_step(idx[k].cuda() if host_indices else None, seed=42, step=s)
        cur = 1e-3
        for m in ms:
            if m <= s:
                cur *= 0.2
        b.opt_step("Adamax", cur)
        torch.cuda.synchronize()
        if s < ms[0]:  # same learning rate so far: the same loss and parameters to the bit
            np.testing.assert_array_equal(loss[k].numpy(), want.cpu().numpy())
            for j in range(n):
                np.testing.assert_array_equal(a.get_params(j), b.get_params(j))
        else:
            np.testing.assert_allclose(loss[k].numpy(), want.cpu().numpy(), rtol=1e-5)
    for j in range(n):
        np.testing.assert_allclose(a.get_params(j), b.get_params(j), rtol=1e-4, atol=1e-7)
    assert a.steps_done == steps
    if not host_indices:
        c, keep_c = _small_group(prec, n, batch=batch)
        hist = c.fit_run(steps, "Adamax", 1e-3, milestones=ms, gamma=0.2, seed=42, loss_history=True)
        torch.cuda.synchronize()
        for j in range(n):
            np.testing.assert_array_equal(a.get_params(j), c.get_params(j))
        np.testing.assert_array_equal(hist[-1].cpu().numpy(), loss[(steps - 1) & 1].numpy())
        c.close()
    a.close(); b.close()


def test_block_histogram_is_exact():
    from brief_pytorch_b200.group import block_histogram
    rng = np.random.default_rng(0)
    for dtype, hi in ((np.uint16, 65536), (np.uint8, 256)):
        for n in (1, 31, 32, 1000, 300007):
            x = (rng.gamma(1.5, hi / 50, n).clip(0, hi - 1)).astype(dtype)
            x[: n // 3] = 7  # a heavy bin (warp-aggregated path)
            t = torch.from_numpy(x.view(np.int16) if dtype == np.uint16 else x).cuda()
            np.testing.assert_array_equal(block_histogram(t, np.dtype(dtype).name), np.bincount(x, minlength=hi))


def test_quantile_rule_limits_from_the_device_histogram():
    from brief_pytorch_b200 import misc, synth
    blk = synth.neuron((20, 30, 25), seed=9)
    t = torch.from_numpy(np.ascontiguousarray(blk[..., 0]).view(np.int16)).cuda()
    rules = ["quantile_1000_0.25_0.9_0.3", "value_60000_65535_0.5"]
    assert misc.weight_rules_for_kernel(t, rules, np.uint16) == misc.weight_rules_for_kernel(blk, rules)
    # and the weights the kernel forms from those limits are the reference's parse_weight
    from brief_pytorch_b200.group import NetSpec, SirenGroup
    grp = SirenGroup([NetSpec(8, 3, 10.0, blk.shape[:3])], 0, "fp32")
    grp.bind_volume(0, t, float(blk.min()), float(blk.max()), 0.0, 100.0, rules=misc.weight_rules_for_kernel(t, rules, np.uint16),
                    np_dtype="uint16")
    idx = torch.arange(blk.size, dtype=torch.int64, device="cuda")
    _, _, w = grp.gather(0, idx)
    np.testing.assert_array_equal(w.cpu().numpy().ravel(), O.parse_weight(blk.copy(), rules).ravel())
    grp.close()


def test_compress_divide_preprocesses_the_whole_volume_before_dividing(tmp_path):
    """main.py:518-531: threshold + binary opening run on the WHOLE volume, then the partition and the by_var budgets are
    taken from the preprocessed data; the blocks are fitted with denoise off.  A dark run that straddles a block seam is
    zeroed by the whole-volume opening but would survive a per-block one."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"].update(divide_type="total_1_2_2", param_alloc="by_var")
    o["Compress"]["param"]["filesize_ratio"] = 8
    o["Compress"]["checkpoints"] = "none"
    o["Compress"]["preprocess"] = {"denoise": {"level": 300, "close": [2, 2, 2]}, "clip": [0, 60000]}
    vol = synth.vessel((16, 48, 48), seed=7)
    vol[4:6, 23:25, 10:12] = 5   # 2x2x2 dark cube across the h seam (rows 23 | 24)
    vol[4:6, 30:32, 23:25] = 5   # and across the w seam
    cf = NFGR(o, 0, "f16")
    blocks, mine = cf.compress_divide(vol.copy(), str(tmp_path / "c"), max_steps=3)
    pre = O.preprocess(vol.copy(), 300, [2, 2, 2], [0, 60000])
    assert (pre[4:6, 23:25, 10:12] == 0).all() and (pre[4:6, 30:32, 23:25] == 0).all()
    chunks = O.alloc_param(O.divide_data(pre, "total_1_2_2"), vol.nbytes / 8, "by_var", 26)
    assert [b.name for b in blocks] == [c["name"] for c in chunks]
    for b, c in zip(blocks, chunks):
        np.testing.assert_array_equal(b.dev.cpu().numpy().view(np.uint16), c["data"][..., 0])
        np.testing.assert_array_equal(b.data, c["data"])
        np.testing.assert_allclose(b.param_size, c["param_size"], rtol=1e-12)
        assert b.features == O.calc_features(c["param_size"] / 4.0, 3, 1, 7)
        assert b.sideinfos["min"] == float(c["data"].min()) and b.sideinfos["max"] == float(c["data"].max())
    with open(tmp_path / "c" / "sideinfos.yaml") as fh:
        assert yaml.safe_load(fh)["chunks_numbers"] == 4


def test_constant_block_does_not_abort_the_group(tmp_path):
    """An all-background block (max == min) is 0/0 in the reference's normalisation; here it must not take the other
    blocks down, and it decodes to its constant."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    vol = synth.vessel((16, 48, 48), seed=7)
    vol[:, :24, :24] = 777
    cf = NFGR(o, 0, "f16")
    cdir = str(tmp_path / "c")
    blocks, _ = cf.compress_divide(vol, cdir, max_steps=40)
    assert all(np.isfinite(b.loss) for b in blocks)
    assert blocks[0].sideinfos["min"] == blocks[0].sideinfos["max"] == 777.0
    out = cf.decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"), os.path.join(cdir, "sideinfos"))
    assert (out[:, :24, :24] == 777).all()
    assert O.cal_psnr(vol, out, 65535) > 20


def test_chunks_numbers_counts_the_partition_not_the_survivors(tmp_path):
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"].update(divide_type="total_1_2_2", param_alloc="by_var", param_size_thres=600)
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    vol = synth.vessel((16, 48, 48), seed=7)
    vol[:, :24, :24] = (vol[:, :24, :24] // 64) + 100   # a nearly flat block: its by_var share falls below the threshold
    cdir = str(tmp_path / "c")
    blocks, mine = NFGR(o, 0, "f16").compress_divide(vol, cdir, max_steps=2)
    assert len(blocks) == 3
    with open(os.path.join(cdir, "sideinfos.yaml")) as fh:
        assert yaml.safe_load(fh)["chunks_numbers"] == 4   # main.py:529 records it before alloc_param drops blocks
    # rank 1 of 2 writes its blocks but not the shared top-level file
    c2 = str(tmp_path / "c2")
    NFGR(o, 0, "f16").compress_divide(vol, c2, max_steps=2, rank=1, world=2)
    assert not os.path.exists(os.path.join(c2, "sideinfos.yaml")) and os.path.isdir(os.path.join(c2, "module"))


def test_divide_exception_overrides_one_blocks_configuration():
    """Compress.divide.exception (main.py:535-537, 568-569): the named chunk's task config is merged over the default.
    Here one block gets lr 0 (its parameters stay at their initial values) and another a deeper network."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    from brief_pytorch_b200.group import pack_module_params
    from brief_pytorch_b200.Networks import init_phi
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    frozen, deep = "d_0_15-h_0_23-w_24_47", "d_0_15-h_24_47-w_0_23"
    o["Compress"]["divide"]["exception"] = {
        frozen: {"CompressFramework": {"Compress": {"lr_phi": 0.0}}},
        deep: {"CompressFramework": {"Module": {"phi": {"layers": 5}}}}}
    vol = synth.vessel((16, 48, 48), seed=7)
    blocks, _ = NFGR(o, 0, "f16").compress_divide(vol, None, max_steps=30)
    by_name = {b.name: b for b in blocks}
    torch.manual_seed(42)
    init = init_phi(dict(o["Module"]["phi"], features=by_name[frozen].features))
    np.testing.assert_array_equal(pack_module_params(by_name[frozen].module), pack_module_params(init))
    assert len(by_name[deep].module.net) == 5 and len(by_name[frozen].module.net) == 7
    assert by_name[deep].features == O.calc_features(by_name[deep].param_size / 4.0, 3, 1, 5)
    other = [b for b in blocks if b.name not in (frozen, deep)]
    for b in other:
        torch.manual_seed(42)
        p0 = pack_module_params(init_phi(dict(o["Module"]["phi"], features=b.features)))
        assert np.abs(pack_module_params(b.module) - p0).max() > 1e-4


def test_single_task_compress_entry_writes_the_reference_layout(tmp_path):
    """NFGR.compress(data_path) (main.py:322-454): checkpoints -> steps{N}/compressed/{sideinfos.yaml, module/*}, decoded
    quality rows in performance.csv; the oracle decodes the written module to the volume our decode gives."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = opt()
    o["Compress"]["divide"]["divide_type"] = "none"
    o["Compress"]["param"]["filesize_ratio"] = 40
    o["Compress"]["max_steps"] = 40
    o["Compress"]["checkpoints"] = "every_20"
    o["Compress"]["decompress"] = True
    o["Decompress"].update(keep_decompressed=True, mse=True, psnr=True, ssim=True)
    vol = synth.vessel((16, 40, 40), seed=11)
    path = str(tmp_path / "vessel.npy")
    np.save(path, vol)
    logdir = NFGR(o, 0, "f16").compress(path, str(tmp_path / "run"))
    rows = open(os.path.join(logdir, "performance.csv")).read().strip().splitlines()
    assert rows[0].split(",")[:4] == ["steps", "mse", "psnr", "ssim"] and len(rows) == 3
    for step in (20, 40):
        comp = os.path.join(logdir, f"steps{step}", "compressed")
        side = yaml.safe_load(open(os.path.join(comp, "sideinfos.yaml")))
        assert side["data_shape"] == [16, 40, 40, 1] and side["dtype"] == "uint16"
        assert len(os.listdir(os.path.join(comp, "module"))) == 14
    m = O.init_phi(dict(o["Module"]["phi"], features=side["phi_features"]))
    O.load_model(m, os.path.join(comp, "module"))
    ref = O.decompress_block(m, side, "minmaxany_0_100")
    ours = np.load(os.path.join(logdir, "steps40", "decompressed", "vessel_decompressed.npy"))
    span = float(vol.max()) - float(vol.min())
    assert np.abs(ours.astype(np.int64) - ref.astype(np.int64)).max() <= np.ceil(3e-2 * span) + 1
    psnr40 = float(rows[2].split(",")[2])
    assert abs(psnr40 - O.cal_psnr(vol, ours, 65535)) < 0.05
