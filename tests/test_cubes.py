"""General RandomCubeSampler (main.py:38-125: cube_len smaller than the block, cube_count > 1) and lr schedules with more
than 8 decays — CPU side: the oracle against the golden vectors written from the unmodified reference
(oracle/gen_golden_cubes.py), against the reference's live class when /root/reference is mounted, the device function's
window -> voxel arithmetic compiled for the host, and the host logic that feeds the kernels (NFGR's sampler choice,
the BriefOptConfig milestone windows).  The GPU half is tests/test_gpu_cubes.py."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import brief_oracle as O
import refshim
from conftest import ROOT, load_gold

KW3 = dict(coords_channel=3, data_channel=1, layers=5, name="SIREN", w0=20, features=24)
KW2 = dict(coords_channel=2, data_channel=1, layers=4, name="SIREN", w0=30, features=16)


def oracle_sampler(g, tag, rules):
    vol = g[f"{tag}_vol"]
    weight = O.parse_weight(vol.copy(), rules)
    data_t, _ = O.normalize_data(vol.copy(), "minmaxany_0_100")
    return vol, data_t, O.RandomCubeSampler(data_t, weight, "-1,1", int(g[f"{tag}_cube_count"]), [int(c) for c in g[f"{tag}_cube_len"]],
                                            len(g[f"{tag}_ids"]))


@pytest.mark.parametrize("tag,kw,rules", [("c3d", KW3, ["value_10001_65535_0.1"]), ("c2d", KW2, ["none"])])
def test_oracle_cube_sampler_replays_the_reference(tag, kw, rules):
    g = load_gold("cubes")
    vol, data_t, s = oracle_sampler(g, tag, rules)
    assert s.pop_size == int(g[f"{tag}_pop"])
    torch.manual_seed(42)
    O.init_phi(dict(kw))                      # the generator's RNG position: seed -> init_phi -> window draws
    flat = data_t.reshape(-1).numpy()
    for step, (c, d, w) in enumerate(s):
        np.testing.assert_array_equal(s.last_idx.numpy(), g[f"{tag}_ids"][step])
        if step == 0:
            for name, t in (("coords0", c), ("data0", d), ("weight0", w)):
                assert t.numpy().tobytes() == g[f"{tag}_{name}"].tobytes()
        vi = O.cube_voxel_indices(vol.shape[:-1], g[f"{tag}_cube_len"], s.last_idx.numpy())
        np.testing.assert_array_equal(flat[vi].reshape(d.shape), d.numpy())


@pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("shape,clen,count", [((9, 7, 8), [4, 3, 5], 3), ((6, 5, 4), [100, 2, 4], 2), ((5, 6, 7), [9, 9, 9], 1),
                                              ((12, 10), [5, 4], 4), ((12, 10), [50, 4, 7], 2)])
def test_oracle_cube_sampler_equals_the_live_reference_class(shape, clen, count):
    """main.py cannot be imported; refshim cuts its two sampler classes out of the syntax tree and runs them unmodified."""
    live = refshim.load_main_samplers()
    vol = np.arange(int(np.prod(shape)), dtype=np.float32).reshape(*shape, 1)   # data == voxel index
    a = live.RandomCubeSampler(torch.from_numpy(vol.copy()), (vol * 0.5).astype(np.float32), "-1,1", count, list(clen), 4)
    b = O.RandomCubeSampler(torch.from_numpy(vol.copy()), (vol * 0.5).astype(np.float32), "-1,1", count, list(clen), 4)
    assert a.pop_size == b.pop_size
    torch.manual_seed(3)
    want = list(a)
    torch.manual_seed(3)
    for (c1, d1, w1), (c2, d2, w2) in zip(want, b):
        assert torch.equal(c1, c2) and torch.equal(d1, d2) and torch.equal(w1, w2)
        vi = O.cube_voxel_indices(shape, clen, b.last_idx.numpy())
        np.testing.assert_array_equal(vi.reshape(d2.shape[:-1]), d2[..., 0].numpy().astype(np.int64))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not found")
def test_device_cube_arithmetic_on_the_host(tmp_path):
    """brief_cube_voxel is __host__ __device__: the function the index kernel calls, compiled for the host, against the
    oracle for 3-D and 2-D (d = 1) blocks, first / last / random windows, clamped and unclamped cube lengths."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "cube_index_host")
    res = subprocess.run([nvcc, "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tests", "cuda", "cube_index_host.cu")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    rng = np.random.default_rng(4)
    for shape, clen in (((9, 7, 8), (4, 3, 5)), ((20, 24, 28), (8, 6, 10)), ((1, 40, 48), (1, 7, 9)), ((6, 5, 4), (6, 2, 4)),
                        ((3, 3, 3), (3, 3, 3)), ((5, 4, 7), (1, 1, 1))):
        pop = int(np.prod([n - c + 1 for n, c in zip(shape, clen)]))
        ids = sorted({0, pop - 1, *rng.integers(0, pop, size=6).tolist()})
        vox = int(np.prod(clen))
        out = subprocess.run([exe, str(shape[1]), str(shape[2]), str(clen[1]), str(clen[2]), str(vox)] + [str(i) for i in ids],
                             capture_output=True, text=True, check=True).stdout
        got = np.array([[int(x) for x in line.split()] for line in out.strip().splitlines()], dtype=np.int64)
        np.testing.assert_array_equal(got, O.cube_voxel_indices(shape, clen, ids))


def c_side_lr(lr0, window, gamma, t):
    """brief_capi.cu's lr_at: float cfg fields widened to double, one multiplication per milestone behind step t."""
    lr = float(np.float32(lr0))
    for m in window:
        if m <= t - 1:
            lr *= float(np.float32(gamma))
    return lr


def test_schedules_with_more_than_eight_decays():
    """StepLR (utils/misc.py:191-192) as generated milestones, served to the C side in windows of 8 (schedule_window):
    the learning rate of every step equals torch's scheduler (the golden run's record and a fresh torch.optim run)."""
    from brief_pytorch_b200 import misc
    from brief_pytorch_b200.group import schedule_window
    g = load_gold("cubes")
    opt = misc.configure_lr_scheduler(misc.configure_optimizer(None, "Adam", 1e-3), {"name": "StepLR", "step_size": 3, "gamma": 0.7})
    ms = opt.milestones_until(30)
    assert ms == tuple(range(3, 30, 3)) and len(ms) == 9
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.SGD([p], lr=1e-3)
    sch = torch.optim.lr_scheduler.StepLR(topt, step_size=3, gamma=0.7)
    steps_done, windows = 0, 0
    while steps_done < 30:                           # SirenGroup.fit_run's walk over the windows
        lr0, window, horizon = schedule_window(opt.lr, ms, opt.gamma, steps_done)
        assert len(window) <= 8
        n = 30 - steps_done if horizon is None else min(30 - steps_done, horizon - steps_done)
        assert n >= 1
        for t in range(steps_done + 1, steps_done + n + 1):
            want = topt.param_groups[0]["lr"]
            assert abs(c_side_lr(lr0, window, opt.gamma, t) - want) < 2e-7 * want
            assert abs(want - g["c2d_lrs"][t - 1]) < 1e-15
            assert abs(opt.lr_at(t) - want) < 1e-12 * want
            topt.step()
            sch.step()
        steps_done += n
        windows += 1
    assert windows == 2
    # up to 8 milestones pass through unchanged (the validated MultiStepLR path)
    assert schedule_window(1e-3, (50000, 60000, 70000), 0.2, 65000) == (1e-3, [50000, 60000, 70000], None)
    opt = misc.configure_lr_scheduler(misc.configure_optimizer(None, "Adamax", 1e-3),
                                      {"name": "MultiStepLR", "milestones": [20, 30], "gamma": 0.2})
    for t in range(1, 41):
        assert abs(opt.lr_at(t) - g["c3d_lrs"][t - 1]) < 1e-15
    with pytest.raises(NotImplementedError):
        misc.configure_lr_scheduler(opt, {"name": "CosineAnnealingLR", "T_max": 10})   # utils/misc.py:195-196


@pytest.mark.parametrize("name", ["Adamax", "Adam"])
def test_cyclic_lr_schedule_is_torchs_own(name):
    """CyclicLR (utils/misc.py:189-190) changes lr and, with cycle_momentum, beta1 every step: FusedOptimizer.per_step
    reads both off torch's scheduler, so they equal what the reference's optimiser holds at every step, in any order."""
    from brief_pytorch_b200 import misc
    kw = {"name": "CyclicLR", "base_lr": 1e-4, "max_lr": 1e-3, "step_size_up": 4}
    o = misc.configure_lr_scheduler(misc.configure_optimizer(None, name, 1e-3), kw)
    got = o.per_step(1, 6) + o.per_step(7, 6)
    p = torch.nn.Parameter(torch.zeros(1))
    t = O.configure_optimizer([p], name, 1e-3)
    sch = O.configure_lr_scheduler(t, kw)
    want = []
    for _ in range(12):
        want.append((t.param_groups[0]["lr"], t.param_groups[0]["betas"][0]))
        t.step()
        sch.step()
    assert got == want and want[0] == (1e-4, 0.9) and want[4] == (1e-3, 0.8)
    assert o.per_step(3, 2) == want[2:4] and o.lr_at(5) == 1e-3            # rewinds
    with pytest.raises(NotImplementedError):                               # torch would switch SGD to momentum SGD
        misc.configure_lr_scheduler(misc.configure_optimizer(None, "SGD", 1e-3), kw)
    sgd = misc.configure_lr_scheduler(misc.configure_optimizer(None, "SGD", 1e-3), dict(kw, cycle_momentum=False))
    assert [lr for lr, _ in sgd.per_step(1, 5)] == [w[0] for w in want[:5]]


def test_nfgr_sampler_choice_follows_main_py():
    """main.py:325-334, 367-369: the cube sampler survives while min(block voxels, configured cube voxels) <= 80^3;
    cube_len is clamped to the block; a step draws cube_count windows."""
    from test_framework import opt as make_opt
    from brief_pytorch_b200.CompressFramework import NFGR
    o = make_opt()
    cf = NFGR(o, 0, "f16")
    assert cf._sampler_name(64 ** 3, (64, 64, 64)) == "randomcube" and cf._cube_config((64, 64, 64)) == (1, [64, 64, 64])
    assert cf._step_batch((64, 64, 64)) == 64 ** 3
    assert cf._sampler_name(96 ** 3, (96, 96, 96)) == "randompoint"
    assert cf._step_batch((96, 96, 96)) == o["Compress"]["sampler"]["sample_size"]
    o["Compress"]["sampler"].update(cube_len=[8, 100, 8], cube_count=3)     # 6400 <= 80^3: sliding cubes on any block size
    cf = NFGR(o, 0, "f16")
    assert cf._sampler_name(96 ** 3, (96, 96, 96)) == "randomcube"
    assert cf._cube_config((96, 50, 96)) == (3, [8, 50, 8]) and cf._step_batch((96, 50, 96)) == 3 * 8 * 50 * 8


def test_half_selects_the_two_byte_width_rule():
    """Compress.half (main.py:217-220, 242-245): the byte budget buys 2-byte parameters, so the same budget gives a wider
    network; widths and theoretical sizes equal the oracle's estimate_module_size(half=True)."""
    from test_framework import opt as make_opt
    from brief_pytorch_b200.CompressFramework import NFGR
    o = make_opt()
    o["Compress"]["half"] = True
    cf, full = NFGR(o), NFGR(make_opt())
    phi = {k: v for k, v in o["Module"]["phi"].items() if k != "name"}
    for budget in (1000.0, 6516.0, 16241 * 4.0, 64976 * 4.0):
        f, nbytes = cf.estimate_module_size(budget)
        fo, po, so = O.estimate_module_size(budget, dict(phi), half=True)
        assert (f, nbytes) == (fo, so) and nbytes == po * 2.0
        assert f > full.estimate_module_size(budget)[0]


def test_schedule_windows_random_milestone_lists():
    """Any MultiStepLR list (duplicates included: torch applies gamma once per occurrence), any resume point: walking
    the run window by window gives the C side the learning rate of torch's scheduler at every step."""
    from brief_pytorch_b200.group import schedule_window
    rng = np.random.default_rng(6)
    for _ in range(60):
        n_steps = int(rng.integers(5, 120))
        ms = sorted(int(m) for m in rng.integers(1, n_steps + 10, size=int(rng.integers(0, 30))))
        gamma = float(rng.choice([0.1, 0.2, 0.5, 0.9]))
        p = torch.nn.Parameter(torch.zeros(1))
        topt = torch.optim.SGD([p], lr=1e-3)
        sch = torch.optim.lr_scheduler.MultiStepLR(topt, milestones=ms, gamma=gamma)
        want = []
        for _ in range(n_steps):
            want.append(topt.param_groups[0]["lr"])
            topt.step()
            sch.step()
        steps_done = int(rng.integers(0, n_steps))      # a resumed run starts anywhere
        while steps_done < n_steps:
            lr0, window, horizon = schedule_window(1e-3, ms, gamma, steps_done)
            assert len(window) <= 8 and (horizon is None or horizon > steps_done)
            n = n_steps - steps_done if horizon is None else min(n_steps - steps_done, horizon - steps_done)
            for t in range(steps_done + 1, steps_done + n + 1):
                # BriefOptConfig carries lr and gamma as C floats: 6e-8 relative per factor
                assert abs(c_side_lr(lr0, window, gamma, t) - want[t - 1]) <= (len(ms) + 2) * 6e-8 * want[t - 1], (ms, gamma, t)
            steps_done += n


def test_quantile_from_histogram_is_numpys_quantile():
    """The device path of the 'quantile' weight rule reads the two order statistics off a histogram: bit-equal to
    np.quantile on the selected voxels for integer data, any threshold and quantile, ties included."""
    from brief_pytorch_b200 import misc
    rng = np.random.default_rng(7)
    for _ in range(300):
        top = int(rng.choice([255, 65535]))
        n = int(rng.integers(1, 400))
        if rng.random() < 0.5:
            data = rng.integers(0, top + 1, size=n)
        else:
            data = rng.choice(rng.integers(0, top + 1, size=int(rng.integers(1, 5))), size=n)
        ge = float(rng.choice([0, int(np.median(data)), int(data.max()), rng.integers(0, top + 1), 0.5 + int(data.min())]))
        q = float(rng.choice([0.0, 1.0, 0.5, np.round(rng.uniform(0, 1), 3)]))
        sel = data[data >= ge]
        got = misc.quantile_from_histogram(np.bincount(data, minlength=top + 1), ge, q)
        if sel.size == 0:
            assert np.isnan(got)
        else:
            assert got == float(np.quantile(sel, q)), (top, n, ge, q)
