"""Parity of the CUDA path (through the C-ABI, libbrief_b200.so) against the CPU oracle and the golden fixtures
made from the unmodified reference.  Tolerances (BASELINE.json north_star): fp32 mode 1e-4 relative, f16
tensor-core mode 1e-2 relative, of the tensor's max magnitude; index / byte / integer work is bit-exact."""
import numpy as np
import pytest
import torch

import brief_oracle as O
from conftest import load_gold, packed_params, unpack

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "f16": 1e-2}
NETS = {"c1": dict(coords_channel=3, layers=5, w0=20, features=22),
        "c2": dict(coords_channel=3, layers=7, w0=10, features=56),
        "c2small": dict(coords_channel=3, layers=7, w0=10, features=13),
        "img2d": dict(coords_channel=2, layers=5, w0=30, features=32)}


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make_group(specs, prec):
    from brief_pytorch_b200.group import SirenGroup
    from brief_pytorch_b200._cabi import BriefError
    try:
        return SirenGroup(specs, 0, prec)
    except BriefError as e:
        if e.code == -3 and prec == "f16":
            pytest.skip("shape outside the tcgen05 kernel: " + str(e))
        raise


def spec_of(kw, dims):
    from brief_pytorch_b200.group import NetSpec
    return NetSpec(kw["features"], kw["layers"], kw["w0"], dims, kw["coords_channel"])


def u16(t):
    return t.cpu().numpy().view(np.uint16)


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("tag", sorted(NETS))
def test_forward_per_layer(tag, prec):
    g, kw = load_gold("siren_" + tag), NETS[tag]
    grp = make_group([spec_of(kw, (4, 4, 4) if kw["coords_channel"] == 3 else (4, 4))], prec)
    assert grp.precision(0) == prec
    grp.set_params(0, packed_params(g, kw["layers"]))
    np.testing.assert_array_equal(grp.get_params(0), packed_params(g, kw["layers"]))
    y, zs = grp.forward(0, torch.from_numpy(g["coords"]).cuda(), return_layers=True)
    for l in range(kw["layers"] - 1):
        assert relerr(zs[l].cpu().numpy(), g[f"z{l}"]) < TOL[prec], f"z{l}"
    assert relerr(y.cpu().numpy(), g["y"]) < TOL[prec]


def bind_block(grp, net, blk, rules=(), tau=0.0, weight=None):
    raw = torch.from_numpy(blk.reshape(blk.shape[:-1]).view(np.int16)).cuda().contiguous()
    w = None if weight is None else torch.from_numpy(weight.reshape(-1).astype(np.float32)).cuda()
    grp.bind_volume(net, raw, float(blk.min()), float(blk.max()), 0.0, 100.0, weight=w, rules=rules, tau=tau,
                    np_dtype="uint16")


def test_gather_is_bit_exact():
    g = load_gold("train_small")
    blk = g["block"]
    kw = dict(coords_channel=3, layers=5, w0=20, features=22)
    grp = make_group([spec_of(kw, blk.shape[:3])], "fp32")
    grp.set_axes(0, "-1,1")
    bind_block(grp, 0, blk, rules=[(10001, 65535, 0.1)])
    idx = torch.from_numpy(g["l5_idx"][0])
    c, d, w = grp.gather(0, idx.cuda())
    data_t, _ = O.normalize_data(blk.copy(), "minmaxany_0_100")
    coords = O.create_flattened_coords(blk.shape[:3], "-1,1")
    assert c.cpu().numpy().tobytes() == coords[idx].numpy().tobytes()
    assert d.cpu().numpy().tobytes() == data_t.reshape(-1, 1)[idx].numpy().tobytes()
    assert w.cpu().numpy().tobytes() == g["l5_weight"].reshape(-1, 1)[idx.numpy()].tobytes()
    # explicit weight volume instead of on-chip rules; whole-block order (RandomCubeSampler shape of output)
    bind_block(grp, 0, blk, weight=g["l5_weight"])
    c, d, w = grp.gather(0, None, blk.size)
    assert c.cpu().numpy().tobytes() == coords.numpy().tobytes()
    assert w.cpu().numpy().tobytes() == g["l5_weight"].reshape(-1, 1).tobytes()


def test_device_sampler_matches_oracle_stream():
    from brief_pytorch_b200.group import sample_indices
    for seed, step, net, batch, pop in ((42, 0, 0, 1000, 262144), (42, 79999, 63, 100000, 16777216), (7, 5, 2, 33, 17)):
        got = sample_indices(seed, step, net, batch, pop).cpu().numpy()
        np.testing.assert_array_equal(got, O.device_sample_indices(seed, step, net, batch, pop))


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("tag", ["l5", "l7"])
def test_loss_and_gradients(tag, prec):
    g = load_gold("train_small")
    layers, w0, f = (int(x) for x in g[f"{tag}_cfg"])
    blk, thr = g["block"], float(g[f"{tag}_thr"])
    kw = dict(coords_channel=3, layers=layers, w0=w0, features=f)
    grp = make_group([spec_of(kw, blk.shape[:3])], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, g[f"{tag}_Adamax_p0"])
    bind_block(grp, 0, blk, weight=g[f"{tag}_weight"], tau=thr)
    grp.set_sampler(0, "randompoint", 300)
    loss = grp.fit_step(torch.from_numpy(g[f"{tag}_idx"][0]).cuda())
    assert abs(float(loss[0]) - g[f"{tag}_Adamax_losses"][0]) < TOL[prec] * g[f"{tag}_Adamax_losses"][0]
    grads = unpack(grp.get_grads(0), 3, f, layers)
    for l in range(layers):
        assert relerr(grads[l][0], g[f"{tag}_dW{l}"]) < 5 * TOL[prec], f"dW{l}"
        assert relerr(grads[l][1].reshape(-1), g[f"{tag}_db{l}"]) < 5 * TOL[prec], f"db{l}"


@pytest.mark.parametrize("optname", ["Adamax", "Adam", "SGD"])
@pytest.mark.parametrize("tag", ["l5", "l7"])
def test_optimiser_kernel_on_reference_gradients(tag, optname):
    """Feed the reference's own step-0 gradients to the fused optimiser kernel: p1 must match to fp32 rounding."""
    g = load_gold("train_small")
    layers, w0, f = (int(x) for x in g[f"{tag}_cfg"])
    kw = dict(coords_channel=3, layers=layers, w0=w0, features=f)
    grp = make_group([spec_of(kw, (8, 12, 10))], "fp32")
    grp.set_params(0, g[f"{tag}_{optname}_p0"])
    grp.set_grads(0, np.concatenate([np.concatenate([g[f"{tag}_dW{l}"].ravel(), g[f"{tag}_db{l}"].ravel()])
                                     for l in range(layers)]))
    grp.opt_step(optname, 1e-3, t=1)
    p1 = grp.get_params(0)
    ref = g[f"{tag}_{optname}_p1"]
    assert np.abs(p1 - ref).max() <= 2 * np.spacing(np.abs(ref).max().astype(np.float32))


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("optname", ["Adamax", "Adam", "SGD"])
def test_four_training_steps_replayed_indices(optname, prec):
    g = load_gold("train_small")
    for tag in ("l5", "l7"):
        layers, w0, f = (int(x) for x in g[f"{tag}_cfg"])
        blk, thr = g["block"], float(g[f"{tag}_thr"])
        grp = make_group([spec_of(dict(coords_channel=3, layers=layers, w0=w0, features=f), blk.shape[:3])], prec)
        grp.set_axes(0, "-1,1")
        grp.set_params(0, g[f"{tag}_{optname}_p0"])
        bind_block(grp, 0, blk, weight=g[f"{tag}_weight"], tau=thr)
        grp.set_sampler(0, "randompoint", 300)
        lr = 1e-3
        for step in range(4):
            loss = grp.fit_step(torch.from_numpy(g[f"{tag}_idx"][step]).cuda())
            grp.opt_step(optname, lr)
            if step + 1 in (2, 3):  # MultiStepLR([2,3], 0.2) of the fixture
                lr *= 0.2
            ref_l = g[f"{tag}_{optname}_losses"][step]
            assert abs(float(loss[0]) - ref_l) < 3 * TOL[prec] * ref_l
            ref_p = g[f"{tag}_{optname}_p{step + 1}"]
            # Adam/Adamax normalise the update to ~lr per parameter, so compare on the scale of the update
            assert np.abs(grp.get_params(0) - ref_p).max() < (2e-5 if prec == "fp32" else 2.5e-3)


def test_inverse_normalisation_and_truncating_cast_are_bit_exact():
    """invnormalize_data (utils/io.py:136-147): one 1-voxel network per probe value, output = last-layer bias."""
    g = load_gold("normalize")
    from brief_pytorch_b200.group import NetSpec
    yhat = np.concatenate([g["probe"], g["yhat"][:2039]]).astype(np.float32)
    specs = [NetSpec(1, 2, 1.0, (1, 1, 1)) for _ in yhat]
    for dtype, vmin, vmax, ref in (("uint16", float(g["block"].min()), float(g["block"].max()), g["yhat_inv_u16"]),
                                   ("uint8", 3.0, 250.0, g["yhat_inv_u8"])):
        grp = make_group(specs, "fp32")
        for i, y in enumerate(yhat):
            grp.set_params(i, np.float32([0, 0, 0, 0, 0, y]))
            grp.set_denorm(i, vmin, vmax, 0.0, 100.0)
        out = grp.decompress(dtype)
        got = np.array([int(u16(t).ravel()[0]) if dtype == "uint16" else int(t.cpu().numpy().ravel()[0]) for t in out])
        np.testing.assert_array_equal(got[9:], ref[:2039])
        if dtype == "uint16":
            exp = O.invnormalize_data(torch.from_numpy(g["probe"]).clone(), {"dtype": "uint16", "min": vmin, "max": vmax},
                                      "minmaxany_0_100")
            np.testing.assert_array_equal(got[:9], exp)
        grp.close()


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_config1_decompress_of_reference_parameters(prec):
    g, vol = load_gold("config1_200"), load_gold("brain64")["volume"]
    kw = NETS["c1"]
    grp = make_group([spec_of(kw, (64, 64, 64))], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, g["p_final"])
    grp.set_denorm(0, float(g["vmin"]), float(g["vmax"]))
    rec = grp.decompress("float32")[0].cpu().numpy()
    assert relerr(rec, g["rec_fp32"][..., 0]) < TOL[prec]
    dec = u16(grp.decompress("uint16")[0])
    diff = np.abs(dec.astype(np.int64) - g["decompressed"][..., 0].astype(np.int64))
    counts_per_unit = (float(g["vmax"]) - float(g["vmin"])) / 100.0
    assert diff.max() <= np.ceil(TOL[prec] * 100 * counts_per_unit) + 1
    if prec == "fp32":
        assert (diff > 0).mean() < 0.01
    a = vol.astype(np.float32)
    assert abs(O.cal_psnr(a, dec[..., None].astype(np.float32), 65535) - float(g["psnr"])) < 0.1
    assert abs(O.cal_ssim(a, dec[..., None].astype(np.float32), 65535) - float(g["ssim"])) < 0.002


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_config1_200_steps_match_reference_quality(prec):
    """SingleTask default.yaml: 200 full-batch Adamax steps on the shipped 64^3 block (loop enqueued from C)."""
    g, vol = load_gold("config1_200"), load_gold("brain64")["volume"]
    grp = make_group([spec_of(NETS["c1"], (64, 64, 64))], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, g["p0"])
    bind_block(grp, 0, vol, rules=[(65535, 65535, 1.0)], tau=float(g["thr"]))
    grp.set_sampler(0, "randomcube")
    hist = grp.fit_run(200, "Adamax", 1e-3, milestones=(50000, 60000, 70000), gamma=0.2, loss_history=True)
    losses = hist[:, 0].cpu().numpy()
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-3 if prec == "fp32" else 2e-2)
    dec = u16(grp.decompress("uint16")[0])[..., None]
    a = vol.astype(np.float32)
    assert abs(O.cal_psnr(a, dec.astype(np.float32), 65535) - float(g["psnr"])) < 0.1
    assert abs(O.cal_ssim(a, dec.astype(np.float32), 65535) - float(g["ssim"])) < 0.002


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_networks_in_a_group_are_independent(prec):
    """Blocks are independent networks: a network's losses, parameters and decoded block are bit-identical
    whether it is fitted alone or next to others (the property multi-GPU sharding relies on)."""
    g = load_gold("train_small")
    blk = g["block"]
    kws = [dict(coords_channel=3, layers=5, w0=20, features=22), dict(coords_channel=3, layers=7, w0=10, features=24),
           dict(coords_channel=3, layers=7, w0=10, features=13)]
    p0 = [g["l5_Adamax_p0"], g["l7_Adamax_p0"]]
    torch.manual_seed(1)
    from brief_pytorch_b200 import Networks
    from brief_pytorch_b200.group import pack_module_params
    p0.append(pack_module_params(Networks.init_phi(dict(kws[2], name="SIREN"))))

    def run(members):
        grp = make_group([spec_of(kws[i], blk.shape[:3]) for i in members], prec)
        for j, i in enumerate(members):
            grp.set_axes(j, "-1,1")
            grp.set_params(j, p0[i])
            bind_block(grp, j, blk, rules=[(10001, 65535, 0.1)], tau=50.0)
            grp.set_sampler(j, "randomcube" if i != 1 else "randompoint", 700)
        idx = torch.arange(700, dtype=torch.int64).cuda() % blk.size
        out = []
        for step in range(3):
            loss = grp.fit_step(idx if 1 in members else None)
            grp.opt_step("Adamax", 1e-3)
            out.append(loss.cpu().numpy())
        dec = [u16(t) for t in grp.decompress("uint16")]
        return np.stack(out), [grp.get_params(j) for j in range(len(members))], dec

    l_all, p_all, d_all = run([0, 1, 2])
    for i in range(3):
        l_one, p_one, d_one = run([i])
        assert l_one[:, 0].tobytes() == l_all[:, i].tobytes()
        assert p_one[0].tobytes() == p_all[i].tobytes()
        assert d_one[0].tobytes() == d_all[i].tobytes()


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_full_size_block_properties(prec):
    """At BASELINE size (vessel block 64x256x256, L=7 f=56, batch 100000): decompress is deterministic, agrees
    with forward() on the same coordinates, and the on-device sampler's fit lowers the loss."""
    from brief_pytorch_b200 import Networks
    from brief_pytorch_b200.group import pack_module_params
    kw = NETS["c2"]
    dims = (64, 256, 256)
    torch.manual_seed(42)
    phi = Networks.init_phi(dict(kw, data_channel=1, name="SIREN"))
    grp = make_group([spec_of(kw, dims)], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, pack_module_params(phi))
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, n, dtype=np.float32) for n in dims], indexing="ij")
    vol = (20000 + 9000 * np.sin(4 * zz + 2 * yy) * np.cos(3 * xx)).astype(np.uint16)[..., None]
    bind_block(grp, 0, vol, rules=[(65535, 65535, 1.0)], tau=0.0)
    grp.set_sampler(0, "randompoint", 100000)
    a = grp.decompress("float32")[0]
    b = grp.decompress("float32")[0]
    assert torch.equal(a, b)
    coords = O.create_flattened_coords(dims, "-1,1")[1234567:1234567 + 4096].cuda()
    y = grp.forward(0, coords)
    assert torch.equal(y.reshape(-1), a.reshape(-1)[1234567:1234567 + 4096])
    hist = grp.fit_run(60, "Adamax", 1e-3, seed=42, loss_history=True)[:, 0].cpu().numpy()
    assert np.isfinite(hist).all() and hist[-10:].mean() < hist[:10].mean()
    dec = u16(grp.decompress("uint16")[0])
    assert dec.min() >= vol.min() and dec.max() <= vol.max()


@pytest.mark.parametrize("optname", ["Adamax", "SGD"])
@pytest.mark.parametrize("tag", ["c1", "c2", "c2small", "img2d"])
def test_operand_image_refreshed_by_the_optimiser_equals_a_full_pack(tag, optname):
    """The optimiser kernel rewrites the fp16 operand image entry by entry (image_scatter); a group that gets the
    same parameters through set_params() builds the image with pack_kernel.  Both must decode to identical bits."""
    from brief_pytorch_b200 import Networks
    from brief_pytorch_b200.group import pack_module_params
    kw = NETS[tag]
    dims = (6, 20, 24) if kw["coords_channel"] == 3 else (40, 36)
    torch.manual_seed(3)
    phi = Networks.init_phi(dict(kw, data_channel=1, name="SIREN"))
    rng = np.random.default_rng(5)
    vol = rng.integers(100, 30000, size=dims, dtype=np.uint16)[..., None]
    grp = make_group([spec_of(kw, dims)], "f16")
    grp.set_axes(0, "-1,1")
    grp.set_params(0, pack_module_params(phi))
    bind_block(grp, 0, vol, rules=[(65535, 65535, 1.0)], tau=0.0)
    grp.set_sampler(0, "randomcube")
    grp.fit_run(7, optname, 1e-3 if optname == "Adamax" else 1e-6, seed=1)   # image kept current by the optimiser
    a = grp.decompress("float32")[0].cpu().numpy()
    fresh = make_group([spec_of(kw, dims)], "f16")
    fresh.set_axes(0, "-1,1")
    fresh.set_params(0, grp.get_params(0))                                    # image rebuilt by pack_kernel
    b = fresh.decompress("float32")[0].cpu().numpy()
    assert a.tobytes() == b.tobytes()
    assert np.isfinite(a).all()


@pytest.mark.parametrize("features", [70, 90, 113, 126, 140, 180, 228, 254])
def test_wide_networks_decode_on_the_tensor_core(features):
    """Widths above the narrow fit kernel's envelope (hipct f=113): under BRIEF_PREC_AUTO the fit runs the
    wide tcgen05 kernel (streamed weights, stashed activations) and forward / decompress run the tcgen05 decode kernel
    (operand image kept current by the optimiser).  Per-layer pre-activations and the decoded block against the oracle,
    f16 tolerance."""
    from brief_pytorch_b200 import Networks
    from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
    kw = dict(coords_channel=3, data_channel=1, layers=7, w0=10, features=features)
    dims = (5, 24, 40)
    torch.manual_seed(11)
    phi = Networks.init_phi(dict(kw, name="SIREN"))
    torch.manual_seed(11)
    ora = O.init_phi(dict(kw, name="SIREN"))
    grp = SirenGroup([NetSpec(features, 7, 10.0, dims)], 0, "auto")
    assert grp.precision(0) == "f16"           # the fit path: tc_fit_wide_kernel (f <= 126) / the layer-wise kernels (f <= 254)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, pack_module_params(phi))
    coords = O.create_flattened_coords(dims, "-1,1")
    y, zs = grp.forward(0, coords.cuda(), return_layers=True)
    with torch.no_grad():
        y_ref, z_ref, _ = O.forward_layers(O.siren_params(ora), coords, 10.0)
    for l in range(6):
        assert relerr(zs[l].cpu().numpy(), z_ref[l].numpy()) < TOL["f16"], f"z{l}"
    assert relerr(y.cpu().numpy(), y_ref.numpy()) < TOL["f16"]
    # a few optimiser steps must leave the tensor-core image in sync with the parameters
    rng = np.random.default_rng(2)
    vol = rng.integers(1000, 30000, size=dims, dtype=np.uint16)[..., None]
    bind_block(grp, 0, vol, rules=[(65535, 65535, 1.0)], tau=0.0)
    grp.set_sampler(0, "randomcube")
    grp.fit_run(3, "Adamax", 1e-3, seed=1)
    a = grp.decompress("float32")[0].cpu().numpy()
    fresh = SirenGroup([NetSpec(features, 7, 10.0, dims)], 0, "auto")
    fresh.set_axes(0, "-1,1")
    fresh.set_params(0, grp.get_params(0))
    b = fresh.decompress("float32")[0].cpu().numpy()
    assert a.tobytes() == b.tobytes()
    grp.store_module(0, ora)
    with torch.no_grad():
        y2 = O.forward_layers(O.siren_params(ora), coords, 10.0)[0].numpy().reshape(dims)
    assert relerr(a, y2) < TOL["f16"]


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("features,layers,sampler", [(113, 7, "randompoint"), (70, 7, "randomcube"), (90, 5, "randompoint"),
                                                     (126, 7, "randomcube"), (100, 3, "randomcube"),
                                                     (140, 7, "randompoint"), (180, 7, "randomcube"), (228, 7, "randompoint"),
                                                     (228, 5, "randomcube"), (254, 3, "randomcube"), (127, 4, "randomcube")])
def test_wide_networks_loss_and_gradients(features, layers, sampler, prec):
    """The wide fit kernel (64 < F_PAD <= 128), the layer-wise tensor-core kernels (128 < F_PAD <= 256: neuron.yaml as
    shipped, f = 228) and the fp32 kernels at the same widths against the oracle's autograd on the same samples: loss and
    every gradient tensor, several slices and a ragged last tile; then 20 optimiser steps stay on the oracle's loss curve."""
    if prec == "fp32" and features > 126:
        pytest.skip("fp32 CUDA-core kernels at these widths are covered up to f = 126; the subject here is the tensor-core path")
    from brief_pytorch_b200 import Networks
    from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
    kw = dict(coords_channel=3, data_channel=1, layers=layers, w0=10, features=features)
    dims = (6, 21, 37)   # 4662 voxels: 36 full tiles + one of 54 rows
    torch.manual_seed(5)
    phi = Networks.init_phi(dict(kw, name="SIREN"))
    grp = SirenGroup([NetSpec(features, layers, 10.0, dims)], 0, prec)
    assert grp.precision(0) == prec
    grp.set_axes(0, "-1,1")
    grp.set_params(0, pack_module_params(phi))
    rng = np.random.default_rng(4)
    vol = rng.integers(500, 30000, size=dims, dtype=np.uint16)[..., None]
    bind_block(grp, 0, vol, rules=[(10001, 65535, 0.1)], tau=0.0)
    n_vox = int(np.prod(dims))
    if sampler == "randompoint":
        idx = torch.from_numpy(rng.integers(0, n_vox, 3001)).long()
        grp.set_sampler(0, "randompoint", 3001)
    else:
        idx = torch.arange(n_vox)
        grp.set_sampler(0, "randomcube")
    loss = grp.fit_step(idx.cuda() if sampler == "randompoint" else None)
    # oracle: the same samples through torch autograd (fp32, CPU)
    data_t, side = O.normalize_data(vol.copy(), "minmaxany_0_100")
    weight = torch.from_numpy(O.parse_weight(vol, ["value_10001_65535_0.1"]))
    coords = O.create_flattened_coords(dims, "-1,1")[idx]
    y = data_t.reshape(-1, 1)[idx]
    w = weight.reshape(-1, 1)[idx]
    torch.manual_seed(5)
    ora = O.init_phi(dict(kw, name="SIREN"))
    ref_loss, _, ref_grads, _ = O.loss_and_grads(O.siren_params(ora), coords, y, w, 0.0, 10.0)
    assert abs(float(loss[0]) - float(ref_loss)) < TOL[prec] * float(ref_loss)
    grads = unpack(grp.get_grads(0), 3, features, layers)
    for l in range(layers):
        assert relerr(grads[l][0], ref_grads[l][0].numpy()) < 5 * TOL[prec], f"dW{l}"
        assert relerr(grads[l][1].reshape(-1), ref_grads[l][1].numpy().reshape(-1)) < 5 * TOL[prec], f"db{l}"
    # a short fit follows the oracle's loss curve (full-batch cube only: the same samples every step)
    if sampler == "randomcube":
        hist = grp.fit_run(20, "Adamax", 1e-3, seed=1, loss_history=True).cpu().numpy()[:, 0]
        opt = torch.optim.Adamax(ora.parameters(), lr=1e-3)
        ref_hist = []
        for _ in range(20):
            opt.zero_grad()
            l_ = O.datal2(y, ora(coords), w.clone(), 0.0)
            l_.backward()
            opt.step()
            ref_hist.append(float(l_.detach()))
        assert np.abs(hist - np.array(ref_hist)).max() < 3 * TOL[prec] * ref_hist[0]


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_config1_3000_steps_quality_tracks_the_reference(prec):
    """Mid-fit horizon (oracle/gen_golden_long.py: the unmodified reference, 3000 full-batch Adamax steps on the shipped
    block, PSNR 41.91 dB / SSIM 0.9893).  The loss still falls 1.5 % per 10 steps here, so two runs that differ only in
    fp32 summation order sit a few steps apart (measured: fp32 kernels -0.10 dB, f16 kernels -0.05 dB); the bars at this
    horizon are 0.25 dB / 0.004 — the north-star bars (0.1 dB / 0.002) are asserted at 200 steps above and at the
    config's full 20000-step budget below."""
    g, g0, vol = load_gold("config1_3000"), load_gold("config1_200"), load_gold("brain64")["volume"]
    grp = make_group([spec_of(NETS["c1"], (64, 64, 64))], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, g0["p0"])
    bind_block(grp, 0, vol, rules=[(65535, 65535, 1.0)], tau=float(g0["thr"]))
    grp.set_sampler(0, "randomcube")
    hist = grp.fit_run(int(g["steps"]), "Adamax", 1e-3, milestones=(50000, 60000, 70000), gamma=0.2, loss_history=True)
    losses = hist[:, 0].cpu().numpy()[249::250]
    # trajectories of a fit that is still descending fast (1.5 % per 10 steps here) drift apart through rounding alone:
    # the loss is compared loosely, the decoded quality at the north-star bars
    np.testing.assert_allclose(losses, g["losses"], rtol=5e-2)
    dec = u16(grp.decompress("uint16")[0])[..., None]
    a = vol.astype(np.float32)
    psnr, ssim = O.cal_psnr(a, dec.astype(np.float32), 65535), O.cal_ssim(a, dec.astype(np.float32), 65535)
    print(f"[{prec}] psnr {psnr:.4f} (reference {float(g['psnr']):.4f})  ssim {ssim:.5f} (reference {float(g['ssim']):.5f})")
    assert abs(psnr - float(g["psnr"])) < 0.25
    assert abs(ssim - float(g["ssim"])) < 0.004


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_wide_network_600_steps_track_the_reference(prec):
    """A wide network (L = 5, f = 70: F_PAD = 80, the wide tcgen05 fit kernel's first bucket) fitted for 600 full-batch
    Adamax steps on the shipped block against the UNMODIFIED reference on the CPU (oracle/gen_golden_wide.py): the loss
    curve at every step of the first 100 and every 50th after, and the decoded PSNR / SSIM at the end."""
    g, vol = load_gold("wide70_600"), load_gold("brain64")["volume"]
    f, steps = int(g["features"]), int(g["steps"])
    kw = dict(coords_channel=3, layers=5, w0=20, features=f)
    grp = make_group([spec_of(kw, (64, 64, 64))], prec)
    assert grp.precision(0) == prec
    grp.set_axes(0, "-1,1")
    grp.set_params(0, g["p0"])
    bind_block(grp, 0, vol, rules=[(65535, 65535, 1.0)], tau=float(g["thr"]))
    grp.set_sampler(0, "randomcube")
    hist = grp.fit_run(steps, "Adamax", 1e-3, milestones=(50000, 60000, 70000), gamma=0.2, loss_history=True)
    losses = hist[:, 0].cpu().numpy()
    ref = g["losses"]
    np.testing.assert_allclose(losses[:100], ref[:100], rtol=3 * TOL[prec])
    np.testing.assert_allclose(losses[49::50], ref[49::50], rtol=5e-2)   # a fast-descending fit drifts through rounding alone
    dec = u16(grp.decompress("uint16")[0])[..., None]
    a = vol.astype(np.float32)
    psnr, ssim = O.cal_psnr(a, dec.astype(np.float32), 65535), O.cal_ssim(a, dec.astype(np.float32), 65535)
    print(f"[{prec}] psnr {psnr:.4f} (reference {float(g['psnr']):.4f})  ssim {ssim:.5f} (reference {float(g['ssim']):.5f})")
    assert abs(psnr - float(g["psnr"])) < 0.25
    assert abs(ssim - float(g["ssim"])) < 0.004


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_config1_full_budget_quality_within_north_star_bars(prec):
    """SingleTask default.yaml's full step budget (20000 full-batch Adamax steps on the shipped block), golden from the
    unmodified reference on the CPU (oracle/gen_golden_long.py 20000): decoded PSNR within 0.1 dB and SSIM within 0.002
    of the reference at equal steps (BASELINE.json north_star)."""
    import os
    from conftest import GOLD
    if not os.path.exists(os.path.join(GOLD, "config1_20000.npz")):
        pytest.skip("tests/golden/config1_20000.npz not generated (oracle/gen_golden_long.py 20000, ~3 h of CPU)")
    g, g0, vol = load_gold("config1_20000"), load_gold("config1_200"), load_gold("brain64")["volume"]
    grp = make_group([spec_of(NETS["c1"], (64, 64, 64))], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, g0["p0"])
    bind_block(grp, 0, vol, rules=[(65535, 65535, 1.0)], tau=float(g0["thr"]))
    grp.set_sampler(0, "randomcube")
    grp.fit_run(int(g["steps"]), "Adamax", 1e-3, milestones=(50000, 60000, 70000), gamma=0.2)
    dec = u16(grp.decompress("uint16")[0])[..., None]
    a = vol.astype(np.float32)
    psnr, ssim = O.cal_psnr(a, dec.astype(np.float32), 65535), O.cal_ssim(a, dec.astype(np.float32), 65535)
    print(f"[{prec}] psnr {psnr:.4f} (reference {float(g['psnr']):.4f})  ssim {ssim:.5f} (reference {float(g['ssim']):.5f})")
    assert abs(psnr - float(g["psnr"])) < 0.1
    assert abs(ssim - float(g["ssim"])) < 0.002
