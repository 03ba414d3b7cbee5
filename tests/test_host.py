"""Host-side logic of brief_pytorch_b200 (no GPU): the C-ABI library loads and exports every symbol the header
declares, the Python mirror of the reference boundary reproduces the golden fixtures, and compute entry points
fail loudly without a CUDA device (there is no CPU fallback)."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest
import torch

import brief_oracle as O
from conftest import ROOT, load_gold, packed_params

from brief_pytorch_b200 import _cabi
from brief_pytorch_b200 import ModelSave, Networks, dataset, io as bio, misc


def header_symbols():
    text = open(os.path.join(ROOT, "include", "brief_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(brief_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_cabi._build.build())
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/brief_b200.h but not exported"
    assert sorted(_cabi.PROTOTYPES) == names, "ctypes prototypes and header disagree"
    assert _cabi.load().brief_abi_version() == 1


def test_linspace_matches_torch_bit_for_bit():
    lib = _cabi.load()
    for n in (1, 2, 3, 5, 7, 16, 33, 64, 100, 256, 511, 1024):
        out = (ctypes.c_float * n)()
        assert lib.brief_linspace(-1.0, 1.0, n, out) == 0
        assert np.frombuffer(out, np.float32).tobytes() == torch.linspace(-1, 1, n).numpy().tobytes(), n
    assert lib.brief_linspace(-1.0, 1.0, 0, None) < 0
    assert b"n=0" in lib.brief_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    lib = _cabi.load()
    desc = (_cabi.NetDesc * 1)(_cabi.NetDesc(3, 1, 22, 5, 20.0, 30.0, (ctypes.c_int32 * 3)(4, 4, 4)))
    h = ctypes.c_void_p()
    rc = lib.brief_group_create(desc, 1, 0, _cabi.PREC_AUTO, ctypes.byref(h))
    assert rc == -2 and not h.value  # BRIEF_ERR_CUDA
    from brief_pytorch_b200.group import NetSpec, SirenGroup
    with pytest.raises(_cabi.BriefError):
        SirenGroup([NetSpec(22, 5, 20.0, (4, 4, 4))])
    phi = Networks.init_phi({"name": "SIREN", "features": 8, "layers": 3})
    with pytest.raises(RuntimeError, match="no CPU compute path"):
        phi(torch.zeros(4, 3))


@pytest.mark.parametrize("tag,kw", [("c1", dict(coords_channel=3, layers=5, w0=20, features=22)),
                                    ("c2", dict(coords_channel=3, layers=7, w0=10, features=56)),
                                    ("img2d", dict(coords_channel=2, layers=5, w0=30, features=32))])
def test_constructor_is_bit_identical_to_reference(tag, kw):
    g = load_gold("siren_" + tag)
    torch.manual_seed(42)
    phi = Networks.init_phi(dict(kw, data_channel=1, name="SIREN", output_act=False, res=False))
    assert (torch.randint(0, 262144, (5,)).numpy() == g["next_randint"]).all()
    for l in range(kw["layers"]):
        assert phi.net[l][0].weight.detach().numpy().tobytes() == g[f"W{l}"].tobytes()
        assert phi.net[l][0].bias.detach().numpy().tobytes() == g[f"b{l}"].tobytes()
    assert Networks.get_nnmodule_param_count(phi) == Networks.SIREN.calc_param_count(data_channel=1, **kw)
    from brief_pytorch_b200.group import pack_module_params
    assert pack_module_params(phi).tobytes() == packed_params(g, kw["layers"]).tobytes()


def test_constructor_errors_and_width_solver():
    with pytest.raises(KeyError):
        Networks.init_phi({"name": "NeRF"})
    with pytest.raises(NotImplementedError):
        Networks.SIREN(res=True)
    for layers, budget, f, p in load_gold("features")["rows"]:
        kw = dict(coords_channel=3, data_channel=1, layers=int(layers))
        assert Networks.SIREN.calc_features(param_count=budget / 4.0, **kw) == int(f)
        assert Networks.SIREN.calc_param_count(features=int(f), **kw) == int(p)


def test_model_files_are_byte_identical_to_reference_layout():
    g = load_gold("config1_200")
    phi = Networks.init_phi({"name": "SIREN", "features": 22, "layers": 5, "w0": 20})
    from brief_pytorch_b200.group import unpack_module_params
    unpack_module_params(phi, g["p_final"])
    ora = O.init_phi({"name": "SIREN", "features": 22, "layers": 5, "w0": 20})
    with tempfile.TemporaryDirectory() as td:
        a, b = os.path.join(td, "mine"), os.path.join(td, "oracle")
        ModelSave.save_model(phi, a)
        assert sorted(os.listdir(a)) == list(g["module_files"])
        O.load_model(ora, a)
        O.save_model(ora, b)
        for f in g["module_files"]:
            assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read()
        phi2 = Networks.init_phi({"name": "SIREN", "features": 22, "layers": 5, "w0": 20})
        ModelSave.load_model(phi2, b)
        from brief_pytorch_b200.group import pack_module_params
        assert pack_module_params(phi2).tobytes() == g["p_final"].astype(np.float32).tobytes()
        ModelSave.save_model(phi2, a)  # replaces the directory like the reference
        assert len(os.listdir(a)) == 10


def test_coords_normalise_weights_checkpoints():
    g = load_gold("coords")
    for key in g.files:
        if key.startswith("axis_"):
            _, n, mode = key.split("_")
            assert dataset.axis_table(int(n), mode).numpy().tobytes() == g[key].tobytes()
        else:
            shp = tuple(int(s) for s in key[5:].split("x"))
            assert dataset.create_flattened_coords(shp, "-1,1").numpy().tobytes() == g[key].tobytes()
    n = load_gold("normalize")
    t, side = bio.normalize_data(n["block"].copy(), "minmaxany_0_100")
    assert t.numpy().tobytes() == n["normalized"].tobytes()
    np.testing.assert_array_equal(bio.invnormalize_data(torch.from_numpy(n["probe"]), {"dtype": "uint16", "min": 16633.0,
                                  "max": 24070.0}, "minmaxany_0_100"), n["probe_inv"])
    assert bio.normalized_threshold(65535, "minmaxany_0_100", 16633.0, 24070.0) == float(n["thres_norm"])
    rules = (["value_65535_65535_1"], ["value_10001_65535_0.1"], ["value_0_2000_0.5", "value_10001_65535_0.1"],
             ["none"], ["quantile_1000_0.2_0.9_0.3"])
    for i, r in enumerate(rules):
        np.testing.assert_array_equal(misc.parse_weight(n["block"].copy(), r), n[f"w{i}"])
    for cp in ("none", "every_2000", "every_7000", "100,300,90000"):
        assert misc.parse_checkpoints(cp, 20000) == O.parse_checkpoints(cp, 20000)
    with pytest.raises(NotImplementedError):
        bio.normalize_data(n["block"], "zscore")


def test_partition_merge_metrics():
    g = load_gold("partition")
    for d, h, w, nb, ps, nd, nh, nw in g["divnum"]:
        assert list(misc.cal_divide_num(int(d), int(h), int(w), int(nb), float(ps))) == [int(nd), int(nh), int(nw)]
    vol = g["volume"]
    for dt in ("total_2_2_3", "every_5_8_7"):
        chunks, _ = misc.divide_data(vol.copy(), dt)
        assert [c["name"] for c in chunks] == list(g[f"{dt}_names"])
        for alloc in ("equal", "by_size", "by_var"):
            kept = misc.alloc_param([dict(c) for c in chunks], 9000.0, alloc, 26)
            np.testing.assert_array_equal([float(c["param_size"]) for c in kept], g[f"{dt}_{alloc}_sizes"])
        np.testing.assert_array_equal(misc.merge_divided_data(chunks, vol.shape), vol)
    c1, vol = load_gold("config1_200"), load_gold("brain64")["volume"]
    a, b = vol.astype(np.float32), c1["decompressed"].astype(np.float32)
    assert abs(misc.cal_psnr(a, b, 65535) - float(c1["psnr"])) < 1e-9
    assert abs(misc.cal_ssim(a, b, 65535) - float(c1["ssim"])) < 1e-5


def test_optimizer_config_mirror():
    opt = misc.configure_optimizer(None, "Adamax", 1e-3)
    misc.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": [50000, 60000, 70000], "gamma": 0.2})
    assert opt.lr_at(1) == 1e-3 and opt.lr_at(50000) == 1e-3
    assert abs(opt.lr_at(50001) - 2e-4) < 1e-12 and abs(opt.lr_at(70001) - 8e-6) < 1e-12
    with pytest.raises(NotImplementedError):
        misc.configure_optimizer(None, "RMSprop", 1e-3)


def test_alloc_param_by_d_by_dv_match_reference():
    """utils/misc.py:410-422 + utils/adaptive_blocking.py:16-24: golden vectors from the unmodified reference
    (oracle/gen_golden_alloc.py); numpy's FFT on the host like the reference, so the budgets are bit-identical."""
    g = load_gold("alloc_d")
    for tag in ("small", "hipct"):
        vol = g[f"{tag}_volume"]
        for dt in ("total_2_2_2", "every_7_9_11"):
            chunks, _ = misc.divide_data(vol.copy(), dt)
            np.testing.assert_array_equal([misc.cal_feature(c["data"]) for c in chunks], g[f"{tag}_{dt}_feature"])
            for alloc in ("by_d", "by_dv"):
                kept = misc.alloc_param([dict(c) for c in chunks], 9000.0, alloc, 26)
                assert [c["name"] for c in kept] == list(g[f"{tag}_{dt}_{alloc}_names"])
                np.testing.assert_array_equal([float(c["param_size"]) for c in kept], g[f"{tag}_{dt}_{alloc}_sizes"])


def test_alloc_param_by_var_from_block_sums():
    """The device statistics kernel returns exact integer sums; the variance formed from them gives the reference's
    by_var budgets to ~1e-12 relative and the same widths."""
    g = load_gold("partition")
    vol = g["volume"]
    chunks, _ = misc.divide_data(vol.copy(), "total_2_2_3")
    want = g["total_2_2_3_by_var_sizes"]
    for c in chunks:
        x = c["data"].astype(np.int64)
        c["var"] = misc.variance_from_sums(float(x.sum()), float((x * x).sum()), c["size"])
    kept = misc.alloc_param(chunks, 9000.0, "by_var", 26)
    np.testing.assert_allclose([c["param_size"] for c in kept], want, rtol=1e-12)


def test_quantile_from_histogram_is_numpy_quantile():
    rng = np.random.default_rng(7)
    for dtype, hi in ((np.uint16, 65536), (np.uint8, 256)):
        for trial in range(6):
            n = int(rng.integers(1, 5000))
            data = (rng.gamma(2.0, hi / 40, n).clip(0, hi - 1)).astype(dtype)
            hist = np.bincount(data.ravel(), minlength=hi)
            for ge in (0.0, float(np.median(data)), 3.5):
                sel = data[data >= ge]
                if sel.size == 0:
                    continue
                for q in (0.0, 0.1, 0.37, 0.5, 0.9, 0.999, 1.0):
                    assert misc.quantile_from_histogram(hist, ge, q) == float(np.quantile(sel, q)), (dtype, n, ge, q)
    # the kernel-rule translation with a host array goes through np.quantile itself
    data = (rng.gamma(2.0, 2000, (6, 7, 8, 1)).clip(0, 65535)).astype(np.uint16)
    rules = misc.weight_rules_for_kernel(data, ["quantile_100_0.2_0.8_0.5", "value_0_50_2"])
    sel = data[data >= 100]
    assert rules == [(float(np.quantile(sel, 0.2)), float(np.quantile(sel, 0.8)), 0.5), (0.0, 50.0, 2.0)]
    w = misc.parse_weight(data, ["quantile_100_0.2_0.8_0.5", "value_0_50_2"])
    ref_w = np.ones(data.shape, np.float32)
    ref_w[(data >= rules[0][0]) & (data <= rules[0][1])] = 0.5
    ref_w[(data >= 0) & (data <= 50)] = 2
    np.testing.assert_array_equal(w, ref_w)
