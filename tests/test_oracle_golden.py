"""The CPU oracle (oracle/brief_oracle.py) against the golden fixtures that oracle/gen_golden.py produced by
running the UNMODIFIED reference.  Bit-exact wherever the arithmetic does not depend on the host's thread count."""
import os
import tempfile

import numpy as np
import pytest
import torch

import brief_oracle as O
from conftest import load_gold, packed_params

NETS = {"c1": dict(coords_channel=3, layers=5, w0=20, features=22),
        "c2": dict(coords_channel=3, layers=7, w0=10, features=56),
        "c2small": dict(coords_channel=3, layers=7, w0=10, features=13),
        "img2d": dict(coords_channel=2, layers=5, w0=30, features=32)}


@pytest.mark.parametrize("tag", sorted(NETS))
def test_init_and_forward_bit_exact(tag):
    g, kw = load_gold("siren_" + tag), NETS[tag]
    torch.manual_seed(42)
    phi = O.init_phi(dict(kw, data_channel=1, name="SIREN", output_act=False, res=False))
    nxt = torch.randint(0, 262144, (5,)).numpy()
    for l in range(kw["layers"]):
        assert phi.net[l][0].weight.detach().numpy().tobytes() == g[f"W{l}"].tobytes()
        assert phi.net[l][0].bias.detach().numpy().tobytes() == g[f"b{l}"].tobytes()
    assert (nxt == g["next_randint"]).all()
    with torch.no_grad():
        y, zs, _ = O.forward_layers(O.siren_params(phi), torch.from_numpy(g["coords"]), kw["w0"])
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=1e-6)
    for l, z in enumerate(zs):
        np.testing.assert_allclose(z.numpy(), g[f"z{l}"], rtol=0, atol=1e-6)


def test_survey_known_answers():
    """The spot values SURVEY.md section 8c quotes from the reference for the config-1 network."""
    g = load_gold("siren_c1")
    np.testing.assert_array_equal(g["W0"][0], np.float32([-0.14512893557548523, 0.28571709990501404, 0.2293398380279541]))
    np.testing.assert_array_equal(g["b0"][:3], np.float32([0.4788495898246765, -0.34219661355018616, -0.3443305492401123]))
    assert str(g["sha256"]) == "d6547ca5737486fc016083ad79ac28b8e711440d481f6403718fccf001e90fc7"
    assert list(g["next_randint"]) == [125541, 189648, 249634, 256353, 9686]


def test_width_solver():
    for layers, budget, f, p in load_gold("features")["rows"]:
        kw = dict(coords_channel=3, data_channel=1, layers=int(layers), name="SIREN")
        assert O.estimate_module_size(float(budget), kw)[:2] == (int(f), int(p))


def test_coords_bit_exact():
    g = load_gold("coords")
    for key in g.files:
        if key.startswith("axis_"):
            _, n, mode = key.split("_")
            assert O.axis_coords(int(n), mode).numpy().tobytes() == g[key].tobytes(), key
        else:
            shp = tuple(int(s) for s in key[5:].split("x"))
            assert O.create_flattened_coords(shp, "-1,1").numpy().tobytes() == g[key].tobytes(), key


def test_normalize_inverse_weights():
    g = load_gold("normalize")
    t, side = O.normalize_data(g["block"].copy(), "minmaxany_0_100")
    assert t.numpy().tobytes() == g["normalized"].tobytes()
    assert (side["min"], side["max"]) == (float(g["vmin"]), float(g["vmax"]))
    inv = O.invnormalize_data(torch.from_numpy(g["probe"]).clone(), {"dtype": "uint16", "min": 16633.0, "max": 24070.0},
                              "minmaxany_0_100")
    np.testing.assert_array_equal(inv, g["probe_inv"])
    np.testing.assert_array_equal(inv[:7], [16633, 20351, 20351, 24069, 24070, 24070, 16633])
    s16 = {"dtype": "uint16", "min": float(g["block"].min()), "max": float(g["block"].max())}
    np.testing.assert_array_equal(O.invnormalize_data(torch.from_numpy(g["yhat"]).clone(), s16, "minmaxany_0_100"),
                                  g["yhat_inv_u16"])
    np.testing.assert_array_equal(O.invnormalize_data(torch.from_numpy(g["yhat"]).clone(),
                                                      {"dtype": "uint8", "min": 3.0, "max": 250.0}, "minmaxany_0_100"),
                                  g["yhat_inv_u8"])
    rules = (["value_65535_65535_1"], ["value_10001_65535_0.1"], ["value_0_2000_0.5", "value_10001_65535_0.1"],
             ["none"], ["quantile_1000_0.2_0.9_0.3"])
    for i, r in enumerate(rules):
        np.testing.assert_array_equal(O.parse_weight(g["block"].copy(), r), g[f"w{i}"])
    assert O.weight_thres_normalized(65535, "minmaxany_0_100", 16633.0, 24070.0) == float(g["thres_norm"])


@pytest.mark.parametrize("tag", ["l5", "l7"])
def test_grads_and_optimiser_steps(tag):
    g = load_gold("train_small")
    layers, w0, f = (int(x) for x in g[f"{tag}_cfg"])
    blk, weight, thr = g["block"], g[f"{tag}_weight"], float(g[f"{tag}_thr"])
    data_t, _ = O.normalize_data(blk.copy(), "minmaxany_0_100")
    coords = O.create_flattened_coords(blk.shape[:3], "-1,1")
    kw = dict(coords_channel=3, data_channel=1, name="SIREN", layers=layers, w0=w0, features=f)
    idx = torch.from_numpy(g[f"{tag}_idx"])
    for optname in ("Adamax", "Adam", "SGD"):
        torch.manual_seed(42)
        phi = O.init_phi(kw)
        opt = O.configure_optimizer(phi.parameters(), optname, 1e-3)
        sch = O.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": [2, 3], "gamma": 0.2})
        p = np.concatenate([q.detach().numpy().ravel() for q in phi.parameters()])
        assert p.tobytes() == g[f"{tag}_{optname}_p0"].tobytes()
        for step in range(4):
            c = coords[idx[step]]
            d = data_t.reshape(-1, 1)[idx[step]]
            w = torch.from_numpy(weight).reshape(-1, 1)[idx[step]].clone()
            if step == 0 and optname == "Adamax":
                _, yhat, grads, _ = O.loss_and_grads(O.siren_params(phi), c, d, w.clone(), thr, w0)
                np.testing.assert_allclose(yhat.numpy(), g[f"{tag}_yhat0"], atol=1e-6)
                for l in range(layers):
                    np.testing.assert_allclose(grads[l][0].numpy(), g[f"{tag}_dW{l}"], rtol=1e-5, atol=1e-7)
                    np.testing.assert_allclose(grads[l][1].numpy(), g[f"{tag}_db{l}"], rtol=1e-5, atol=1e-7)
            loss = O.train_step(phi, opt, sch, c, d, w, thr)
            np.testing.assert_allclose(float(loss), g[f"{tag}_{optname}_losses"][step], rtol=1e-6)
            p = np.concatenate([q.detach().numpy().ravel() for q in phi.parameters()])
            np.testing.assert_allclose(p, g[f"{tag}_{optname}_p{step + 1}"], rtol=1e-5, atol=1e-8)


def test_partition_helpers():
    g = load_gold("partition")
    for d, h, w, nb, ps, nd, nh, nw in g["divnum"]:
        assert list(O.cal_divide_num(int(d), int(h), int(w), int(nb), float(ps))) == [int(nd), int(nh), int(nw)]
    vol = g["volume"]
    for dt in ("total_2_2_3", "every_5_8_7"):
        chunks = O.divide_data(vol.copy(), dt)
        assert [c["name"] for c in chunks] == list(g[f"{dt}_names"])
        for alloc in ("equal", "by_size", "by_var"):
            kept = O.alloc_param([dict(c) for c in chunks], 9000.0, alloc, 26)
            np.testing.assert_array_equal([float(c["param_size"]) for c in kept], g[f"{dt}_{alloc}_sizes"])
        np.testing.assert_array_equal(O.merge_divided_data(chunks, vol.shape), vol)


def test_config1_first_steps_and_quality():
    """Config 1 (shipped 64^3 brain block, L=5 f=22): the first full-batch Adamax steps reproduce the reference's
    losses; the fixture's 200-step parameters decode to the fixture's volume, PSNR and SSIM."""
    g, vol = load_gold("config1_200"), load_gold("brain64")["volume"]
    data_t, side = O.normalize_data(vol.copy(), "minmaxany_0_100")
    thr = O.weight_thres_normalized(65535, "minmaxany_0_100", side["min"], side["max"])
    assert thr == float(g["thr"])
    kw = dict(coords_channel=3, data_channel=1, name="SIREN", layers=5, w0=20, features=int(g["features"]))
    torch.manual_seed(42)
    phi = O.init_phi(kw)
    assert np.concatenate([q.detach().numpy().ravel() for q in phi.parameters()]).tobytes() == g["p0"].tobytes()
    opt = O.configure_optimizer(phi.parameters(), "Adamax", 1e-3)
    sch = O.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": [50000, 60000, 70000], "gamma": 0.2})
    weight = O.parse_weight(vol.copy(), ["value_65535_65535_1"])
    sampler = O.RandomCubeSampler(data_t, weight, "-1,1", 1, [10000000] * 3, 5)
    losses = [float(O.train_step(phi, opt, sch, c, d, w, thr)) for c, d, w in sampler]
    np.testing.assert_allclose(losses, g["losses"][:5], rtol=1e-5)
    assert abs(g["losses"][0] - 2488.302) < 1e-2 and abs(g["losses"][-1] - 2086.219) < 1e-2
    # decode the fixture's final parameters
    off = 0
    with torch.no_grad():
        for q in phi.parameters():
            q.copy_(torch.from_numpy(g["p_final"][off:off + q.numel()].reshape(tuple(q.shape))))
            off += q.numel()
    side = dict(side, data_shape=list(vol.shape))
    dec = O.decompress_block(phi, side, "minmaxany_0_100")
    assert (dec.astype(np.int64) - g["decompressed"].astype(np.int64)).__abs__().max() <= 1
    a, b = vol.astype(np.float32), dec.astype(np.float32)
    assert abs(O.cal_psnr(a, b, 65535) - float(g["psnr"])) < 1e-3
    assert abs(O.cal_ssim(a, b, 65535) - float(g["ssim"])) < 1e-4
    with tempfile.TemporaryDirectory() as td:
        O.save_model(phi, os.path.join(td, "module"))
        assert sorted(os.listdir(os.path.join(td, "module"))) == list(g["module_files"])
        assert sum(os.path.getsize(os.path.join(td, "module", f)) for f in g["module_files"]) == int(g["module_bytes"]) == 6516


def test_device_sampler_stream_properties():
    idx = O.device_sample_indices(42, 7, 3, 100000, 262144)
    assert idx.min() >= 0 and idx.max() < 262144 and idx.dtype == np.int64
    assert abs(idx.mean() / 262144 - 0.5) < 0.01
    assert not np.array_equal(idx, O.device_sample_indices(42, 8, 3, 100000, 262144))
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors: zero counter/key and the 'pi' vector)
    z = O.philox4x32_10(np.zeros((1, 4), np.uint32), (0, 0))[0]
    assert [int(x) for x in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    p = O.philox4x32_10(np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32), (0xa4093822, 0x299f31d0))[0]
    assert [int(x) for x in p] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
