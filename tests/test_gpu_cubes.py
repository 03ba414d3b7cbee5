"""Reference options no shipped config uses, on the GPU: CyclicLR / StepLR schedules, Compress.half, and above all
the general RandomCubeSampler (main.py:38-125: windows smaller than the block, several per step), through the
C-ABI (brief_group_set_cube_sampler / brief_cube_indices), against tests/golden/cubes.npz — written from the unmodified
reference by oracle/gen_golden_cubes.py — and against the oracle.  Index work is bit-exact; losses are within the
north-star tolerances (fp32 mode 1e-4, f16 mode 1e-2, per step, with headroom for the drift of 30-40 optimiser steps)."""
import copy
import os

import numpy as np
import pytest
import torch

import brief_oracle as O
from conftest import load_gold
from test_framework import opt as vessel_opt
from test_gpu_parity import make_group

pytestmark = pytest.mark.gpu

CASES = {
    "c3d": dict(kw=dict(coords_channel=3, data_channel=1, layers=5, name="SIREN", w0=20, features=24), rules=[(10001, 65535, 0.1)],
                opt="Adamax", np_dtype="uint16"),
    "c2d": dict(kw=dict(coords_channel=2, data_channel=1, layers=4, name="SIREN", w0=30, features=16), rules=[],
                opt="Adam", np_dtype="uint8"),
}


def cube_group(g, tag, prec, params=None):
    """One network on the golden block of `tag`, bound like main.py:336-383 does, with the golden run's cube sampler."""
    from brief_pytorch_b200.group import NetSpec
    case, vol = CASES[tag], g[f"{tag}_vol"]
    kw = case["kw"]
    dims = tuple(int(x) for x in vol.shape[:-1])
    grp = make_group([NetSpec(kw["features"], kw["layers"], float(kw["w0"]), dims, kw["coords_channel"])], prec)
    grp.set_axes(0, "-1,1")
    grp.set_params(0, g[f"{tag}_p0"] if params is None else params)
    a = np.ascontiguousarray(vol[..., 0])
    raw = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda()
    grp.bind_volume(0, raw, float(g[f"{tag}_vmin"]), float(g[f"{tag}_vmax"]), 0.0, 100.0, rules=case["rules"],
                    tau=float(g[f"{tag}_tau"]), np_dtype=case["np_dtype"])
    grp.set_cube_sampler(0, int(g[f"{tag}_cube_count"]), [int(c) for c in g[f"{tag}_cube_len"]])
    return grp, dims


@pytest.mark.parametrize("tag", sorted(CASES))
def test_cube_sampler_outputs_are_the_references(tag):
    """The sampler protocol (main.py:104-125) on replayed torch.randint draws: voxel indices equal the oracle's expansion
    of the golden window draws, and the three tensors of the first step equal the reference's bytes."""
    from brief_pytorch_b200.sampler import RandomCubeSampler
    g = load_gold("cubes")
    grp, dims = cube_group(g, tag, "fp32")
    count, clen = int(g[f"{tag}_cube_count"]), [int(c) for c in g[f"{tag}_cube_len"]]
    s = RandomCubeSampler(grp, 0, 3, count, list(clen), generator="torch")
    assert s.pop_size == int(g[f"{tag}_pop"]) and len(s) == 3
    assert grp.batch(0) == count * int(np.prod(clen))
    torch.manual_seed(42)
    O.init_phi(dict(CASES[tag]["kw"]))        # the golden run's RNG position: seed -> init_phi -> window draws
    for step, (c, d, w) in enumerate(s):
        want = O.cube_voxel_indices(dims, clen, g[f"{tag}_ids"][step]).reshape(-1)
        np.testing.assert_array_equal(s.last_idx.cpu().numpy(), want)
        assert tuple(d.shape) == (count, *clen, 1) and tuple(c.shape) == (count, *clen, len(dims))
        if step == 0:
            for name, t in (("coords0", c), ("data0", d), ("weight0", w)):
                assert t.cpu().numpy().tobytes() == g[f"{tag}_{name}"].tobytes(), name
    assert step == 2
    grp.close()


@pytest.mark.parametrize("prec", ["fp32", "f16"])
@pytest.mark.parametrize("tag", sorted(CASES))
def test_cube_fit_replays_the_reference(tag, prec):
    """Every step of the golden run (the reference's own windows, its learning rate of that step — MultiStepLR for c3d,
    a StepLR with nine decays for c2d): the loss of each step against the reference's."""
    g = load_gold("cubes")
    grp, _ = cube_group(g, tag, prec)
    losses = []
    for step, ids in enumerate(g[f"{tag}_ids"]):
        idx = grp.cube_indices(0, torch.from_numpy(ids))
        losses.append(float(grp.fit_step(idx)[0]))
        grp.opt_step(CASES[tag]["opt"], float(g[f"{tag}_lrs"][step]))
    ref = g[f"{tag}_losses"]
    assert abs(losses[0] - ref[0]) < 3 * (1e-4 if prec == "fp32" else 1e-2) * ref[0]
    np.testing.assert_allclose(losses, ref, rtol=2e-3 if prec == "fp32" else 3e-2)
    grp.close()


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_cube_fit_run_draws_the_device_stream(prec):
    """With the on-device sampler the window draws are the network's Philox stream (draw c of step s = cube c): the
    indices equal the oracle's restatement, and brief_fit_run (index buffer generated inside the step) leaves exactly
    the losses and parameters of explicit brief_cube_indices -> brief_fit_step -> brief_opt_step."""
    g = load_gold("cubes")
    a, dims = cube_group(g, "c3d", prec)
    b, _ = cube_group(g, "c3d", prec)
    count, clen, pop = int(g["c3d_cube_count"]), [int(c) for c in g["c3d_cube_len"]], int(g["c3d_pop"])
    steps = 5
    hist = a.fit_run(steps, "Adamax", 1e-3, seed=42, loss_history=True)
    for s in range(steps):
        idx = b.cube_indices(0, None, seed=42, step=s)
        draws = O.device_sample_indices(42, s, 0, count, pop)
        np.testing.assert_array_equal(idx.cpu().numpy(), O.cube_voxel_indices(dims, clen, draws).reshape(-1))
        loss = b.fit_step(idx)
        b.opt_step("Adamax", 1e-3)
        np.testing.assert_array_equal(hist[s].cpu().numpy(), loss.cpu().numpy())
    np.testing.assert_array_equal(a.get_params(0), b.get_params(0))
    a.close(); b.close()


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_cube_and_point_networks_share_a_group(prec):
    """A group that holds a cube network runs every step from a generated index buffer.  A random-point network next to
    it must see exactly the stream it draws on chip when it is alone (same loss history and parameters, per-network
    slicing), through brief_fit_run and through the graphed host step (step counter read from device memory)."""
    from brief_pytorch_b200 import Networks, synth
    from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
    dims, f, L, batch, steps = (24, 40, 40), 24, 5, 1500, 4
    blk = synth.neuron(dims, seed=3)
    raw = torch.from_numpy(np.ascontiguousarray(blk[..., 0]).view(np.int16)).cuda()
    torch.manual_seed(7)
    p0 = pack_module_params(Networks.init_phi(dict(name="SIREN", layers=L, w0=10, features=f)))

    def build(members):
        grp = SirenGroup([NetSpec(f, L, 10.0, dims) for _ in members], 0, prec)
        grp.set_slicing(True)
        for j, kind in enumerate(members):
            grp.set_params(j, p0)
            grp.bind_volume(j, raw, float(blk.min()), float(blk.max()), 0.0, 100.0, rules=[(10001, 65535, 0.1)], tau=40.0,
                            np_dtype="uint16")
            if kind == "cube":
                grp.set_cube_sampler(j, 3, [6, 10, 8])
            else:
                grp.set_sampler(j, "randompoint", batch)
            grp.set_stream(j, 10 + (0 if kind == "cube" else 1))
        return grp

    mixed, alone, hosted = build(["cube", "point"]), build(["point"]), build(["cube", "point"])
    h_mixed = mixed.fit_run(steps, "Adamax", 1e-3, seed=42, loss_history=True).cpu().numpy()
    h_alone = alone.fit_run(steps, "Adamax", 1e-3, seed=42, loss_history=True).cpu().numpy()
    np.testing.assert_array_equal(h_mixed[:, 1], h_alone[:, 0])
    np.testing.assert_array_equal(mixed.get_params(1), alone.get_params(0))
    assert np.isfinite(h_mixed).all() and (h_mixed[:, 0] != h_mixed[:, 1]).all()
    loss = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    for s in range(steps):
        hosted.fit_step_host(None, loss[s & 1], "Adamax", 1e-3, seed=42)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(loss[s & 1].numpy(), h_mixed[s])
    for j in range(2):
        np.testing.assert_array_equal(hosted.get_params(j), mixed.get_params(j))
    for grp in (mixed, alone, hosted):
        grp.close()


def test_divided_volume_with_sliding_cubes(tmp_path):
    """NFGR.compress_divide with a cube sampler smaller than the blocks (sampler.cube_len / cube_count of the yaml).  A
    step's loss is that of its four windows and swings by an order of magnitude from step to step, so the fit is judged
    on the WHOLE block: the weighted loss of the fitted parameters (evaluated by the oracle) must have fallen from the
    initial parameters' by about as much as in the oracle's own 200-step run with its own window stream — statistical
    parity, as for the on-device point sampler.  The written directory decodes to the volume's shape and dtype."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR
    from brief_pytorch_b200.group import pack_module_params, unpack_module_params
    steps, count, clen = 200, 4, [8, 12, 12]
    o = vessel_opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    o["Compress"]["sampler"].update(cube_len=list(clen), cube_count=count)
    vol = synth.vessel((16, 48, 48), seed=7)
    cdir = str(tmp_path / "compressed")
    cf = NFGR(o, 0, "auto")
    blocks, mine = cf.compress_divide(vol, cdir, max_steps=steps)
    assert mine == list(range(4))
    for b in blocks:
        blk = vol[b.d[0]:b.d[1] + 1, b.h[0]:b.h[1] + 1, b.w[0]:b.w[1] + 1]
        weight = O.parse_weight(blk.copy(), o["Compress"]["loss"]["weight"])
        data_t, side = O.normalize_data(blk.copy(), "minmaxany_0_100")
        thr = O.weight_thres_normalized(65535, "minmaxany_0_100", side["min"], side["max"])
        coords = O.create_flattened_coords(blk.shape[:3], "-1,1")

        def block_loss(phi):
            with torch.no_grad():
                return float(O.datal2(data_t.reshape(-1, 1), phi(coords), torch.from_numpy(weight).reshape(-1, 1).clone(), thr))

        torch.manual_seed(42)
        ora = O.init_phi(dict(o["Module"]["phi"], features=b.features))
        before = block_loss(ora)
        topt = O.configure_optimizer(ora.parameters(), "Adamax", 1e-3)
        sch = O.configure_lr_scheduler(topt, {"name": "none"})
        for c, d, w in O.RandomCubeSampler(data_t, weight, "-1,1", count, list(clen), steps):
            O.train_step(ora, topt, sch, c, d, w, thr)
        drop_ref = before - block_loss(ora)
        unpack_module_params(ora, pack_module_params(b.module))
        drop = before - block_loss(ora)
        assert np.isfinite(b.loss) and drop_ref > 0.02 * before
        assert abs(drop - drop_ref) < 0.5 * drop_ref, (b.name, before, drop, drop_ref)
    out = cf.decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"), os.path.join(cdir, "sideinfos"))
    assert out.shape == vol.shape and out.dtype == vol.dtype


def test_cyclic_lr_through_nfgr_matches_the_oracle():
    """Compress.lr_scheduler_phi = CyclicLR (utils/misc.py:189-190): NFGR.fit_blocks walks torch's own schedule step by
    step (lr and beta1); 20 whole-block Adamax steps against the oracle's loop with the same scheduler, fp32 mode."""
    from brief_pytorch_b200 import synth
    from brief_pytorch_b200.CompressFramework import NFGR, Block
    sched, steps = {"name": "CyclicLR", "base_lr": 1e-4, "max_lr": 1e-2, "step_size_up": 3}, 20
    vol = synth.vessel((16, 24, 24), seed=7)
    final = {}
    for name, sc in (("cyclic", sched), ("none", {"name": "none"})):
        o = vessel_opt()
        o["Compress"]["checkpoints"] = "none"
        o["Compress"]["lr_scheduler_phi"] = sc
        blk = Block("d_0_15-h_0_23-w_0_23", vol, [0, 15], [0, 23], [0, 23], 4.0 * 1700)
        NFGR(o, 0, "fp32").fit_blocks([blk], max_steps=steps).close()
        final[name] = blk.loss
    weight = O.parse_weight(vol.copy(), ["value_65535_65535_1"])
    data_t, side = O.normalize_data(vol.copy(), "minmaxany_0_100")
    thr = O.weight_thres_normalized(65535, "minmaxany_0_100", side["min"], side["max"])
    torch.manual_seed(42)
    phi = O.init_phi(dict(coords_channel=3, data_channel=1, layers=7, name="SIREN", w0=10, features=blk.features))
    topt = O.configure_optimizer(phi.parameters(), "Adamax", 1e-3)
    sch = O.configure_lr_scheduler(topt, sched)
    sampler = O.RandomCubeSampler(data_t, weight, "-1,1", 1, [10000000] * 3, steps)
    for c, d, w in sampler:
        ref = float(O.train_step(phi, topt, sch, c, d, w, thr))
    assert abs(final["cyclic"] - ref) < 2e-3 * ref, (final, ref)
    assert abs(final["none"] - ref) > 0.02 * ref                              # the schedule is in effect (oracle: 4 %)


def test_half_mode_fits_and_decodes(tmp_path):
    """Compress.half: widths from the 2-byte rule, fit and decode through the same grouped kernels; module.half() with
    half coordinates (main.py:388-391) and reconstruct_flattened(half=True) (utils/misc.py:69-84) return fp16 tensors."""
    from brief_pytorch_b200 import misc, synth
    from brief_pytorch_b200.CompressFramework import NFGR
    o = vessel_opt()
    o["Compress"]["divide"]["divide_type"] = "total_1_2_2"
    o["Compress"]["param"]["filesize_ratio"] = 16
    o["Compress"]["checkpoints"] = "none"
    o["Compress"]["half"] = True
    vol = synth.vessel((16, 48, 48), seed=7)
    first, _ = NFGR(copy.deepcopy(o), 0, "auto").compress_divide(vol, None, max_steps=1)
    cdir = str(tmp_path / "compressed")
    cf = NFGR(o, 0, "auto")
    blocks, _ = cf.compress_divide(vol, cdir, max_steps=60)
    phi = {k: v for k, v in o["Module"]["phi"].items() if k != "name"}
    for b0, b in zip(first, blocks):
        assert b.features == O.estimate_module_size(b.param_size, dict(phi), half=True)[0]
        assert np.isfinite(b.loss) and b.loss < b0.loss
    out = cf.decompress_divide(os.path.join(cdir, "sideinfos.yaml"), os.path.join(cdir, "module"), os.path.join(cdir, "sideinfos"))
    assert out.shape == vol.shape and out.dtype == vol.dtype
    m = blocks[0].module.cuda()
    coords = (torch.rand(257, 3, device="cuda") * 2 - 1).half()
    y32 = m(coords.float())
    y16 = copy.deepcopy(m).half()(coords)
    assert y16.dtype == torch.float16 and y16.shape == (257, 1) and y32.dtype == torch.float32
    scale = max(1.0, float(y32.abs().max()))
    assert float((y16.float() - y32).abs().max()) < 0.05 * scale      # fp16-rounded parameters and output
    rec = misc.reconstruct_flattened(list(blocks[0].shape) + [1], 10000, m.forward, half=True)
    assert rec.dtype == torch.float16 and tuple(rec.shape) == tuple(blocks[0].shape) + (1,)
