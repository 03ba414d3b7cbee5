"""Edge cases of the CUDA path against the oracle: raw dtypes other than uint16, blocks smaller than one 128-sample
tile, batches that are not a multiple of the tile, 2-D images, zero-length calls, and the error paths of the C-ABI."""
import numpy as np
import pytest
import torch

import brief_oracle as O
from test_gpu_parity import TOL, relerr, make_group, spec_of

pytestmark = pytest.mark.gpu


def _oracle_step(kw, blk, weight_rules, thr_raw, idx):
    """Loss and gradients of one step on `idx` with the oracle (reference semantics)."""
    torch.manual_seed(9)
    ora = O.init_phi(dict(kw, data_channel=1, name="SIREN"))
    data_t, side = O.normalize_data(blk.copy(), "minmaxany_0_100")
    weight = O.parse_weight(blk.copy(), weight_rules)
    thr = O.weight_thres_normalized(thr_raw, "minmaxany_0_100", side["min"], side["max"]) if thr_raw else 0.0
    coords = O.create_flattened_coords(blk.shape[:-1], "-1,1")
    w = torch.from_numpy(weight).reshape(-1, 1)[idx].clone()
    loss, _, grads, _ = O.loss_and_grads(O.siren_params(ora), coords[idx], data_t.reshape(-1, 1)[idx], w, thr, kw["w0"])
    return ora, side, float(loss), grads, thr


def _bind(grp, blk, side, rules, tau):
    raw = np.ascontiguousarray(blk[..., 0])
    if raw.dtype == np.uint16:
        t = torch.from_numpy(raw.view(np.int16)).cuda()
    else:
        t = torch.from_numpy(raw).cuda()
    grp.bind_volume(0, t, side["min"], side["max"], 0.0, 100.0, rules=rules, tau=tau, np_dtype=raw.dtype.name)
    return t


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("dtype", ["uint8", "float32", "uint16"])
def test_raw_dtypes_fit_and_decode(dtype, prec):
    """uint8 / float32 / uint16 raw blocks: gather is bit-exact, loss and gradients match, decode is within tolerance."""
    from brief_pytorch_b200.group import pack_module_params
    rng = np.random.default_rng(4)
    dims = (6, 18, 22)
    if dtype == "uint8":
        blk = rng.integers(3, 250, size=dims + (1,), dtype=np.uint8)
        rules, rules_txt, thr_raw = [(200, 255, 0.25)], ["value_200_255_0.25"], 255
    elif dtype == "uint16":
        blk = rng.integers(100, 40000, size=dims + (1,), dtype=np.uint16)
        rules, rules_txt, thr_raw = [(10001, 65535, 0.1)], ["value_10001_65535_0.1"], 65535
    else:
        blk = rng.normal(0.0, 3.0, size=dims + (1,)).astype(np.float32)
        rules, rules_txt, thr_raw = [], ["none"], 0
    kw = dict(coords_channel=3, layers=5, w0=20, features=22)
    idx = torch.from_numpy(rng.integers(0, blk.size, size=777))
    ora, side, ref_loss, ref_grads, thr = _oracle_step(kw, blk, rules_txt, thr_raw, idx)
    grp = make_group([spec_of(kw, dims)], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, pack_module_params(ora))
    keep = _bind(grp, blk, side, rules, thr)
    grp.set_sampler(0, "randompoint", 777)
    c, d, w = grp.gather(0, idx.cuda())
    data_t, _ = O.normalize_data(blk.copy(), "minmaxany_0_100")
    assert d.cpu().numpy().tobytes() == data_t.reshape(-1, 1)[idx].numpy().tobytes()
    assert w.cpu().numpy().tobytes() == O.parse_weight(blk.copy(), rules_txt).reshape(-1, 1)[idx.numpy()].tobytes()
    loss = float(grp.fit_step(idx.cuda())[0])
    assert abs(loss - ref_loss) < 3 * TOL[prec] * abs(ref_loss)
    got = grp.get_grads(0)
    ref = np.concatenate([np.concatenate([gw.numpy().ravel(), gb.numpy().ravel()]) for gw, gb in ref_grads])
    assert relerr(got, ref) < 5 * TOL[prec]
    if dtype != "float32":
        dec = grp.decompress(dtype)[0].cpu().numpy()
        dec = dec.view(np.uint16) if dtype == "uint16" else dec
        want = O.decompress_block(ora, dict(side, data_shape=list(blk.shape)), "minmaxany_0_100")[..., 0]
        span = (side["max"] - side["min"]) / 100.0
        assert np.abs(dec.astype(np.int64) - want.astype(np.int64)).max() <= np.ceil(3 * TOL[prec] * 100 * span) + 1
    del keep


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("dims", [(1, 1, 1), (3, 5, 7), (2, 8, 8), (1, 1, 129), (5, 5, 103)])
def test_blocks_around_one_tile(dims, prec):
    """Blocks of 1, 105, 128, 129 and 2575 voxels (below, at and just above the 128-sample tile; ragged tail):
    whole-block fit step and decode against the oracle."""
    from brief_pytorch_b200.group import pack_module_params
    rng = np.random.default_rng(8)
    n = int(np.prod(dims))
    blk = rng.integers(500, 30000, size=dims + (1,), dtype=np.uint16)
    if n == 1:
        blk[...] = 700
    kw = dict(coords_channel=3, layers=5, w0=20, features=13)
    torch.manual_seed(9)
    ora = O.init_phi(dict(kw, data_channel=1, name="SIREN"))
    vmin, vmax = float(blk.min()), float(blk.max()) + (1.0 if n == 1 else 0.0)  # a constant block would normalise to 0/0
    side = {"min": vmin, "max": vmax, "dtype": "uint16"}
    data = ((blk.astype(np.float32) - np.float32(vmin)) / (np.float32(vmax) - np.float32(vmin))) * np.float32(100.0)
    coords = O.create_flattened_coords(dims, "-1,1")
    loss_ref, _, grads_ref, _ = O.loss_and_grads(O.siren_params(ora), coords, torch.from_numpy(data).reshape(-1, 1),
                                                 torch.ones(n, 1), 0.0, kw["w0"])
    grp = make_group([spec_of(kw, dims)], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, pack_module_params(ora))
    keep = _bind(grp, blk, side, [], 0.0)
    grp.set_sampler(0, "randomcube")
    loss = float(grp.fit_step()[0])
    assert abs(loss - float(loss_ref)) < 3 * TOL[prec] * abs(float(loss_ref))
    ref = np.concatenate([np.concatenate([gw.numpy().ravel(), gb.numpy().ravel()]) for gw, gb in grads_ref])
    assert relerr(grp.get_grads(0), ref) < 5 * TOL[prec]
    dec = grp.decompress("float32")[0].cpu().numpy().reshape(-1)
    with torch.no_grad():
        y = O.forward_layers(O.siren_params(ora), coords, kw["w0"])[0].numpy().reshape(-1)
    assert relerr(dec, y) < TOL[prec]
    del keep


def test_zero_length_and_error_paths():
    from brief_pytorch_b200 import _cabi
    from brief_pytorch_b200.group import NetSpec, SirenGroup
    grp = SirenGroup([NetSpec(13, 5, 20.0, (4, 4, 4))], 0, "auto")
    y = grp.forward(0, torch.empty((0, 3), dtype=torch.float32, device="cuda"))   # zero coordinates: a no-op
    assert y.shape == (0, 1)
    with pytest.raises(_cabi.BriefError) as e:      # fit before bind_volume
        grp.fit_step()
    assert e.value.code == -4
    with pytest.raises(_cabi.BriefError):           # network index out of range
        grp.get_params(3)
    with pytest.raises(_cabi.BriefError) as e:      # data_channel 3 is outside the fused kernels
        SirenGroup([NetSpec(13, 5, 20.0, (4, 4, 4), 3, 3)], 0, "auto")
    assert e.value.code == -3
    with pytest.raises(_cabi.BriefError) as e:      # explicit f16 for a width no tensor-core fit kernel covers (F_PAD > 256)
        SirenGroup([NetSpec(300, 7, 10.0, (4, 4, 4))], 0, "f16")
    assert e.value.code == -3
    with pytest.raises(_cabi.BriefError):           # zero networks
        SirenGroup([], 0, "auto")
    # envelope of the tensor-core fit kernels under "auto": fused kernel up to f = 62, wide kernel up to f = 126
    # (F_PAD = f + 2 rounded up to 16 <= 128), layer-wise kernels up to f = 254 (F_PAD <= 256: neuron.yaml as shipped,
    # f = 228), fp32 CUDA-core kernels beyond; the decode follows the same F_PAD
    for f, want in ((62, "f16"), (63, "f16"), (126, "f16"), (127, "f16"), (228, "f16"), (254, "f16")):
        g = SirenGroup([NetSpec(f, 7, 10.0, (4, 4, 4))], 0, "auto")
        assert g.precision(0) == want, (f, g.precision(0))
        g.close()
    with pytest.raises(_cabi.BriefError) as e:      # f = 255 has no tensor-core fit kernel
        SirenGroup([NetSpec(255, 7, 10.0, (4, 4, 4))], 0, "f16")
    assert e.value.code == -3


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("features", [32, 72])
def test_2d_image_fit_matches_oracle(features, prec):
    """coords_channel = 2 (PNG/JPG path of the reference): channel order (h, w), whole-image batch; a width of the fused
    fit kernel and one of the wide kernel."""
    from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
    rng = np.random.default_rng(12)
    dims = (37, 29)
    img = rng.integers(0, 255, size=dims + (1,), dtype=np.uint8)
    kw = dict(coords_channel=2, layers=5, w0=30, features=features)
    torch.manual_seed(9)
    ora = O.init_phi(dict(kw, data_channel=1, name="SIREN"))
    data_t, side = O.normalize_data(img.copy(), "minmaxany_0_100")
    coords = O.create_flattened_coords(dims, "-1,1")
    loss_ref, _, grads_ref, _ = O.loss_and_grads(O.siren_params(ora), coords, data_t.reshape(-1, 1),
                                                 torch.ones(img.size, 1), 0.0, kw["w0"])
    grp = make_group([NetSpec(features, 5, 30.0, dims, 2)], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, pack_module_params(ora))
    t = torch.from_numpy(np.ascontiguousarray(img[..., 0])).cuda()
    grp.bind_volume(0, t, side["min"], side["max"], 0.0, 100.0, np_dtype="uint8")
    grp.set_sampler(0, "randomcube")
    loss = float(grp.fit_step()[0])
    assert abs(loss - float(loss_ref)) < 3 * TOL[prec] * float(loss_ref)
    ref = np.concatenate([np.concatenate([gw.numpy().ravel(), gb.numpy().ravel()]) for gw, gb in grads_ref])
    assert relerr(grp.get_grads(0), ref) < 5 * TOL[prec]


@pytest.mark.parametrize("dtype", ["uint8", "uint16", "float32"])
def test_block_stats_kernel_is_exact(dtype):
    """Device-side min / max (bit-exact) and sum / sum of squares (fp64) of raw blocks, one launch for all blocks;
    ragged sizes and pointers that are not 16-byte aligned."""
    from brief_pytorch_b200.group import block_stats
    rng = np.random.default_rng(21)
    sizes = [1, 7, 15, 16, 17, 4097, 123457, 64 * 64 * 64 + 3]
    arrs, tens = [], []
    for i, n in enumerate(sizes):
        off = i % 5  # misalign the block inside a larger allocation
        if dtype == "uint8":
            a = rng.integers(0, 256, size=n + off, dtype=np.uint8)
            t = torch.from_numpy(a).cuda()[off:]
        elif dtype == "uint16":
            a = rng.integers(0, 65536, size=n + off, dtype=np.uint16)
            t = torch.from_numpy(a.view(np.int16)).cuda()[off:]
        else:
            a = (rng.normal(0, 1000, size=n + off) * rng.choice([1e-3, 1.0, 1e3], size=n + off)).astype(np.float32)
            t = torch.from_numpy(a).cuda()[off:]
        arrs.append(a[off:])
        tens.append(t)
    got = block_stats(tens, dtype)
    assert got.shape == (len(sizes), 4)
    for a, g in zip(arrs, got):
        assert np.float32(g[0]) == np.float32(a.min()) and np.float32(g[1]) == np.float32(a.max())
        a64 = a.astype(np.float64)
        np.testing.assert_allclose(g[2], a64.sum(), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(g[3], (a64 * a64).sum(), rtol=1e-12)
    assert block_stats([]).shape == (0, 4)


@pytest.mark.parametrize("dtype,shape", [("uint16", (5, 37, 53)), ("uint8", (3, 64, 48)), ("uint16", (2, 11, 11))])
def test_quality_kernel_matches_reference_metrics(dtype, shape):
    """mse / psnr / ssim in one launch (brief_volume_quality) against the oracle's restatement of eval_performance
    (utils/misc.py:447-499, utils/ssim.py): PSNR within 1e-3 dB, SSIM within 1e-4 (north-star bars: 0.1 dB / 0.002)."""
    from brief_pytorch_b200 import misc
    rng = np.random.default_rng(17)
    top = 255 if dtype == "uint8" else 65535
    zz, yy, xx = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    a = (0.45 * top * (1 + np.sin(yy / 6.0 + zz) * np.cos(xx / 4.0)) + rng.integers(0, top // 50, size=shape))
    a = np.clip(a, 0, top).astype(dtype)[..., None]
    b = np.clip(a.astype(np.int64) + rng.integers(-top // 40, top // 40 + 1, size=a.shape), 0, top).astype(dtype)
    got = misc.eval_performance(7, a, b, device="cuda")
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    assert abs(got["psnr"] - O.cal_psnr(a32, b32, top)) < 1e-3
    assert abs(got["ssim"] - O.cal_ssim(a32, b32, top)) < 1e-4
    np.testing.assert_allclose(got["mse"], ((a32.astype(np.float64) - b32) ** 2).mean(), rtol=1e-9)
    same = misc.eval_performance(7, a, a, device="cuda")
    assert same["mse"] == 0 and same["psnr"] == float("inf") and abs(same["ssim"] - 1.0) < 1e-6
