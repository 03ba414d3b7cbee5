"""Differential fuzz of the host-side mirrors (brief_pytorch_b200/{misc,io,dataset,Networks,ModelSave}.py) against the
UNMODIFIED reference functions imported from /root/reference through oracle/refshim.py, on seeded random inputs.  Byte and
integer work must be bit-identical; float64 metrics agree to rounding.  Skipped where the reference is not mounted (the GPU
box): the committed golden vectors (tests/test_host.py, tests/test_oracle_golden.py) cover the same functions there.
BRIEF_FUZZ_ROUNDS scales the number of random cases (default 40; 2000 were run when this file was written)."""
import os

import numpy as np
import pytest
import torch

import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not mounted")
ROUNDS = int(os.environ.get("BRIEF_FUZZ_ROUNDS", "40"))


@pytest.fixture(scope="module")
def ref():
    return refshim.load_reference()


def rand_volume(rng, dtype=None, lo=3, hi=14):
    shape = tuple(int(x) for x in rng.integers(lo, hi, size=3))
    dtype = dtype or rng.choice(["uint8", "uint16"])
    top = 255 if dtype == "uint8" else 65535
    kind = rng.integers(0, 3)
    if kind == 0:
        v = rng.integers(0, top + 1, size=shape)
    elif kind == 1:   # few distinct values, ties everywhere
        v = rng.choice(rng.integers(0, top + 1, size=4), size=shape)
    else:             # smooth + noise
        zz, yy, xx = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
        v = np.clip(top * 0.4 * (1 + np.sin(zz / 3.0 + yy / 5.0) * np.cos(xx / 4.0)) + rng.integers(0, max(2, top // 30), size=shape), 0, top)
    return v.astype(dtype)[..., None]


def test_partition_budget_and_merge(ref):
    from brief_pytorch_b200 import misc
    rng = np.random.default_rng(101)
    for _ in range(ROUNDS):
        vol = rand_volume(rng, "uint16", lo=4, hi=13)   # the reference's divide_data paints 2000 into a copy: no uint8
        d, h, w = vol.shape[:3]
        if rng.random() < 0.5:
            n = [int(rng.integers(1, min(4, s) + 1)) for s in (d, h, w)]
            kind = f"total_{n[0]}_{n[1]}_{n[2]}"
        else:
            n = [int(rng.integers(max(1, s // 3), s + 1)) for s in (d, h, w)]
            kind = f"every_{n[0]}_{n[1]}_{n[2]}"
        ours, _ = misc.divide_data(vol.copy(), kind)
        theirs = ref.misc.divide_data(vol.copy(), kind)
        theirs = theirs[0] if isinstance(theirs, tuple) else theirs
        assert [c["name"] for c in ours] == [c["name"] for c in theirs], kind
        for a, b in zip(ours, theirs):
            assert (a["d"], a["h"], a["w"]) == (list(b["d"]), list(b["h"]), list(b["w"])) and a["size"] == b["size"]
            np.testing.assert_array_equal(a["data"], b["data"])
        np.testing.assert_array_equal(misc.merge_divided_data(ours, vol.shape), ref.misc.merge_divided_data(theirs, vol.shape))
        budget, thres = float(rng.integers(200, 200000)), float(rng.choice([0, 26, 400]))
        for alloc in ("equal", "by_size", "by_var", "by_d", "by_dv"):
            if alloc in ("by_d", "by_dv") and min(min(c["data"].shape[:3]) for c in ours) < 2:
                continue
            try:
                want = ref.misc.alloc_param([dict(c) for c in theirs], budget, alloc, thres)
            except Exception as e:   # e.g. all chunks constant: the reference divides by a zero total
                with pytest.raises(type(e)):
                    misc.alloc_param([dict(c) for c in ours], budget, alloc, thres)
                continue
            got = misc.alloc_param([dict(c) for c in ours], budget, alloc, thres)
            assert [c["name"] for c in got] == [c["name"] for c in want], (kind, alloc)
            np.testing.assert_array_equal([float(c["param_size"]) for c in got], [float(c["param_size"]) for c in want])


def test_divide_grid_choice(ref):
    from brief_pytorch_b200 import misc
    rng = np.random.default_rng(102)
    for _ in range(ROUNDS * 3):
        d, h, w = (int(x) for x in rng.choice([16, 24, 48, 64, 96, 100, 128, 250, 256, 512, 1000, 1024], size=3))
        nb = int(rng.choice([-1, 0, 1, 2, 4, 7, 8, 27, 64, 100, 512, 4096]))
        ps = float(rng.choice([1e3, 5444.0, 1e5, 3.3e6, 5e8]))
        assert list(misc.cal_divide_num(d, h, w, nb, ps)) == list(ref.adaptive_blocking.cal_divide_num(d, h, w, nb, ps)), (d, h, w, nb, ps)


def test_weights_checkpoints_normalisation(ref):
    from brief_pytorch_b200 import io as bio, misc
    rng = np.random.default_rng(103)
    for _ in range(ROUNDS):
        vol = rand_volume(rng)
        top = 255 if vol.dtype == np.uint8 else 65535
        rules = []
        for _ in range(int(rng.integers(0, 4))):
            k = rng.integers(0, 4)
            if k == 0:
                lo = int(rng.integers(0, top + 1))
                rules.append(f"value_{lo}_{int(rng.integers(lo, top + 1))}_{rng.choice([0.1, 0.5, 2, 10])}")
            elif k == 1:
                ge = int(np.quantile(vol, rng.uniform(0, 0.6)))
                ql = float(np.round(rng.uniform(0, 0.7), 2))
                rules.append(f"quantile_{ge}_{ql}_{float(np.round(rng.uniform(ql, 1.0), 2))}_{rng.choice([0.3, 3])}")
            elif k == 2:
                rules.append(f"exp_{int(rng.integers(1, top))}_{rng.choice([0.1, 0.5, 0.9])}")
            else:
                rules.append("none")
        got, want = misc.parse_weight(vol.copy(), rules), ref.misc.parse_weight(vol.copy(), rules)
        assert got.dtype == want.dtype and got.tobytes() == want.tobytes(), rules
        # the on-chip form of the same rules (value / quantile -> (lo, hi, scale) triples) gives the same weights
        triples = misc.weight_rules_for_kernel(vol, rules)
        if triples is not None:
            w = np.ones(vol.shape, np.float32)
            for lo, hi, s in triples:
                w[(vol >= lo) & (vol <= hi)] = s
            assert w.tobytes() == np.asarray(want, np.float32).tobytes(), rules
        # normalise / inverse
        name = f"minmaxany_{rng.choice([0, -1, -50])}_{rng.choice([1, 100, 255])}"
        if vol.max() > vol.min():
            t1, s1 = bio.normalize_data(vol.copy(), name)
            t2, s2 = ref.io.normalize_data(vol.copy(), name)
            assert t1.numpy().tobytes() == t2.numpy().tobytes() and s1 == s2
            y = t2 + torch.from_numpy(rng.normal(0, 3, size=tuple(t2.shape)).astype(np.float32))
            np.testing.assert_array_equal(bio.invnormalize_data(y.clone(), s1, name), ref.io.invnormalize_data(y.clone(), s2, name))
            thr = float(rng.integers(0, top + 1))
            want_thr = float(ref.io.normalize_data(np.array(thr), name, max=s2["max"], min=s2["min"])[0])
            assert bio.normalized_threshold(thr, name, s1["min"], s1["max"]) == want_thr
        steps = int(rng.integers(1, 100000))
        for cp in ("none", f"every_{int(rng.integers(1, 9000))}", ",".join(str(int(x)) for x in rng.integers(1, 120000, size=3))):
            assert misc.parse_checkpoints(cp, steps) == ref.misc.parse_checkpoints(cp, steps), (cp, steps)


def test_width_solver_coords_and_metrics(ref):
    from brief_pytorch_b200 import Networks, dataset, misc
    rng = np.random.default_rng(104)
    for _ in range(ROUNDS * 5):
        kw = dict(coords_channel=int(rng.choice([2, 3])), data_channel=int(rng.choice([1, 3])), layers=int(rng.integers(2, 12)),
                  res=bool(rng.integers(0, 2)))
        budget = float(rng.uniform(30, 3e6))
        f = Networks.SIREN.calc_features(param_count=budget, **kw)
        assert f == ref.Networks.SIREN.calc_features(param_count=budget, **kw), (kw, budget)
        f = max(1, f)
        assert Networks.SIREN.calc_param_count(features=f, **kw) == ref.Networks.SIREN.calc_param_count(features=f, **kw)
    for _ in range(ROUNDS):
        shape = tuple(int(x) for x in rng.integers(1, 40, size=int(rng.choice([2, 3]))))
        mode = str(rng.choice(["-1,1", "0,1", "-0.5,0.5", "n11"])) if hasattr(ref.dataset, "create_coords") else "-1,1"
        try:
            want = ref.dataset.create_flattened_coords(shape, mode)
        except Exception as e:
            with pytest.raises(type(e)):
                dataset.create_flattened_coords(shape, mode)
            continue
        assert dataset.create_flattened_coords(shape, mode).numpy().tobytes() == want.numpy().tobytes(), (shape, mode)
        assert dataset.create_coords(shape, mode).numpy().tobytes() == ref.dataset.create_coords(shape, mode).numpy().tobytes()
    for _ in range(max(4, ROUNDS // 5)):
        dtype = str(rng.choice(["uint8", "uint16"]))
        top = 255 if dtype == "uint8" else 65535
        shape = (int(rng.integers(1, 5)), int(rng.integers(11, 40)), int(rng.integers(11, 40)), 1)
        a = rng.integers(0, top + 1, size=shape).astype(dtype)
        b = np.clip(a.astype(np.int64) + rng.integers(-top // 20, top // 20 + 1, size=shape), 0, top).astype(dtype)
        a, b = a.astype(np.float32), b.astype(np.float32)   # eval_performance's casts (utils/misc.py:484-485)
        assert misc.cal_mse(a, b) == ref.misc.cal_mse(a, b)
        assert abs(misc.cal_psnr(a, b, top) - ref.misc.cal_psnr(a, b, top)) < 1e-9
        assert abs(misc.cal_ssim(a, b, top) - ref.misc.cal_ssim(a, b, top)) < 1e-5


def test_model_files_round_trip_both_ways(ref, tmp_path):
    """Files written by either side load into the other with identical parameters and identical bytes on disk."""
    from brief_pytorch_b200 import ModelSave, Networks
    rng = np.random.default_rng(105)
    for i in range(max(3, ROUNDS // 10)):
        kw = dict(coords_channel=int(rng.choice([2, 3])), data_channel=1, layers=int(rng.integers(2, 8)), w0=float(rng.choice([10, 20, 30])),
                  features=int(rng.integers(1, 40)), name="SIREN", output_act=False, res=False)
        seed = int(rng.integers(0, 10000))
        torch.manual_seed(seed)
        ours = Networks.init_phi(dict(kw))
        torch.manual_seed(seed)
        theirs = ref.Networks.init_phi(dict(kw))
        for p, q in zip(ours.parameters(), theirs.parameters()):
            assert p.detach().numpy().tobytes() == q.detach().numpy().tobytes()
        a, b = str(tmp_path / f"ours{i}"), str(tmp_path / f"theirs{i}")
        ModelSave.save_model(ours, a)
        ref.ModelSave.save_model(theirs, b)
        assert sorted(os.listdir(a)) == sorted(os.listdir(b))
        for name in os.listdir(a):
            assert open(os.path.join(a, name), "rb").read() == open(os.path.join(b, name), "rb").read(), name
        torch.manual_seed(seed + 1)
        fresh_ours, fresh_theirs = Networks.init_phi(dict(kw)), ref.Networks.init_phi(dict(kw))
        ModelSave.load_model(fresh_ours, b)
        ref.ModelSave.load_model(fresh_theirs, a)
        for p, q, r in zip(ours.parameters(), fresh_ours.parameters(), fresh_theirs.parameters()):
            assert p.detach().numpy().tobytes() == q.detach().numpy().tobytes() == r.detach().numpy().tobytes()


def test_oracle_training_loop_is_the_references_bit_for_bit(ref):
    """The oracle's restatement of the loop main.py:385-400 (init, sampler, loss, optimiser, scheduler) against the
    reference's own pieces — utils.Networks.init_phi, the live sampler classes and datal2 closure cut out of main.py,
    utils.misc.configure_optimizer / configure_lr_scheduler — on random small configurations: every loss and every
    parameter after four steps identical to the bit.  This is what makes the oracle a stand-in for the reference where
    the reference cannot travel (the GPU box)."""
    import brief_oracle as O
    live = refshim.load_main_samplers()
    rng = np.random.default_rng(106)
    for case in range(max(6, ROUNDS // 4)):
        vol = rand_volume(rng, lo=5, hi=12)
        top = 255 if vol.dtype == np.uint8 else 65535
        if vol.max() == vol.min():
            continue
        kw = dict(coords_channel=3, data_channel=1, layers=int(rng.integers(2, 8)), w0=float(rng.choice([10, 20, 30])),
                  features=int(rng.integers(2, 24)), name="SIREN", output_act=False, res=False)
        optname = str(rng.choice(["Adamax", "Adam", "SGD"]))
        sched = [{"name": "MultiStepLR", "milestones": [1, 3], "gamma": 0.2}, {"name": "StepLR", "step_size": 2, "gamma": 0.5},
                 {"name": "none"}][int(rng.integers(0, 3))]
        rules = [[f"value_{top // 4}_{top}_0.1"], ["none"], [f"exp_{top // 2}_0.5"]][int(rng.integers(0, 3))]
        thres = float(rng.choice([0, top // 3, top]))
        cube = bool(rng.integers(0, 2))
        clen = [int(rng.integers(1, s + 3)) for s in vol.shape[:3]]
        count, batch, seed = int(rng.integers(1, 4)), int(rng.integers(1, 300)), int(rng.integers(0, 1000))

        def run(side):
            N, M, IO, S, loss_fn = ((ref.Networks, ref.misc, ref.io, live, live.datal2) if side == "ref"
                                    else (O, O, O, O, O.datal2))
            weight = M.parse_weight(vol.copy(), rules)
            data_t, info = IO.normalize_data(vol.copy(), "minmaxany_0_100")
            tau = float(IO.normalize_data(np.array(thres), "minmaxany_0_100", max=info["max"], min=info["min"])[0])
            torch.manual_seed(seed)
            phi = N.init_phi(dict(kw))
            opt = M.configure_optimizer(phi.parameters(), optname, 1e-3)
            sch = M.configure_lr_scheduler(opt, sched)
            sampler = (S.RandomCubeSampler(data_t, weight, "-1,1", count, list(clen), 4) if cube
                       else S.RandompointSampler(data_t, weight, "-1,1", batch, 4))
            losses = []
            for c, d, w in sampler:           # main.py:385-400
                opt.zero_grad()
                loss = loss_fn(d, phi.forward(c), w, tau)
                loss.backward()
                opt.step()
                sch.step()
                losses.append(float(loss.detach()))
            return losses, [p.detach().numpy().copy() for p in phi.parameters()]

        l_ref, p_ref = run("ref")
        l_ora, p_ora = run("oracle")
        assert l_ref == l_ora, (case, kw, optname, sched, rules)
        for a, b in zip(p_ref, p_ora):
            assert a.tobytes() == b.tobytes(), (case, kw, optname)


def test_oracle_decode_partition_preprocess_against_reference(ref, tmp_path):
    """The rest of the oracle against the reference on random inputs: dense decode (reconstruct_flattened +
    invnormalize_data), model files, partition / budgets / merge, preprocess (scipy call and the scipy-free restatement)."""
    import brief_oracle as O
    rng = np.random.default_rng(107)
    varied = 0
    for case in range(max(6, ROUNDS // 4)):
        vol = rand_volume(rng, "uint16", lo=4, hi=12)
        if vol.max() == vol.min():
            continue
        kw = dict(coords_channel=3, data_channel=1, layers=int(rng.integers(2, 7)), w0=float(rng.choice([10, 20])),
                  features=int(rng.integers(2, 20)), name="SIREN", output_act=False, res=False)
        seed = int(rng.integers(0, 1000))
        torch.manual_seed(seed)
        phi_r = ref.Networks.init_phi(dict(kw))
        torch.manual_seed(seed)
        phi_o = O.init_phi(dict(kw))
        with torch.no_grad():                 # spread the outputs over the whole 0..100 range and beyond
            for phi in (phi_r, phi_o):
                phi.net[-1][0].weight.mul_(3000.0)
                phi.net[-1][0].bias.fill_(50.0)
        _, side = ref.io.normalize_data(vol.copy(), "minmaxany_0_100")
        side = dict(side, data_shape=list(vol.shape), phi_features=kw["features"], phi_name="SIREN")
        chunk = int(rng.choice([7, 100, 10000]))
        with torch.no_grad():
            rec = ref.misc.reconstruct_flattened(side["data_shape"], chunk, phi_r.forward, device="cpu", coords_mode="-1,1").float().cpu()
        want = ref.io.invnormalize_data(rec.clone(), side, "minmaxany_0_100")
        got = O.decompress_block(phi_o, side, "minmaxany_0_100", sample_size=chunk)
        np.testing.assert_array_equal(got, want)
        varied += len(np.unique(want)) > 2
        # model files: written by one, read by the other
        a, b = str(tmp_path / f"r{case}"), str(tmp_path / f"o{case}")
        ref.ModelSave.save_model(phi_r, a)
        O.save_model(phi_o, b)
        for name in sorted(os.listdir(a)):
            assert open(os.path.join(a, name), "rb").read() == open(os.path.join(b, name), "rb").read()
        # partition / budgets / merge
        n = [int(rng.integers(1, min(3, s) + 1)) for s in vol.shape[:3]]
        kind = f"total_{n[0]}_{n[1]}_{n[2]}"
        theirs = ref.misc.divide_data(vol.copy(), kind)
        theirs = theirs[0] if isinstance(theirs, tuple) else theirs
        ours = O.divide_data(vol.copy(), kind)
        assert [c["name"] for c in ours] == [c["name"] for c in theirs]
        np.testing.assert_array_equal(O.merge_divided_data(ours, vol.shape), ref.misc.merge_divided_data(theirs, vol.shape))
        for alloc in ("equal", "by_size", "by_var"):
            try:
                w_ = ref.misc.alloc_param([dict(c) for c in theirs], 50000.0, alloc, 26.0)
            except Exception:
                continue
            g_ = O.alloc_param([dict(c) for c in ours], 50000.0, alloc, 26.0)
            np.testing.assert_array_equal([float(c["param_size"]) for c in g_], [float(c["param_size"]) for c in w_])
        # preprocess
        level = int(np.quantile(vol, rng.uniform(0.05, 0.6)))
        close = False if rng.random() < 0.25 else [int(x) for x in rng.integers(1, 4, size=3)]
        clip = [int(rng.integers(0, 1000)), int(rng.integers(30000, 65536))]
        want = ref.misc.preprocess(vol.copy(), level, close, clip)
        np.testing.assert_array_equal(O.preprocess(vol.copy(), level, close, clip), want)
        np.testing.assert_array_equal(O.preprocess_restated(vol.copy(), level, close, clip), want)
    assert varied >= 3   # the decodes were not all saturated
