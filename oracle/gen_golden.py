"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
through oracle/refshim.py) and pin the oracle restatement against it.

Run in the build container only:  python oracle/gen_golden.py
Every fixture is produced by the reference's own functions; wherever the oracle restates the same
function the two results are compared BIT-FOR-BIT here and the script aborts on any mismatch, so a
committed fixture certifies "oracle == reference" for that function on that input.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import brief_oracle as O  # noqa: E402
import refshim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
ref = refshim.load_reference()
torch.set_num_threads(max(1, os.cpu_count() or 1))


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    assert a.tobytes() == b.tobytes(), f"oracle != reference for {what}: max|d|={np.abs(a.astype(np.float64)-b.astype(np.float64)).max()}"


def params_np(module):
    out = {}
    for l in range(len(module.net)):
        out[f"W{l}"] = module.net[l][0].weight.detach().numpy().copy()
        out[f"b{l}"] = module.net[l][0].bias.detach().numpy().copy()
    return out


def read_brain():
    import cv2
    path = os.path.join(refshim.REFERENCE_ROOT, "dataset/brain/64x64x64/brain-64_128-64_128-192_256.tif")
    ok, imgs = cv2.imreadmulti(path, flags=cv2.IMREAD_UNCHANGED)
    assert ok
    return np.stack(imgs)[..., None], os.path.getsize(path)


def synth_block(shape, seed, bright_frac=0.02):
    """Small synthetic uint16 block with a dark background and a few bright voxels (neuron-like),
    so that the value_10001_65535_0.1 weight rule has both classes."""
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    base = 3000 + 2500 * np.sin(3 * zz + 1.0) * np.cos(2 * yy) + 1500 * np.sin(5 * xx * yy)
    base += rng.normal(0, 200, shape)
    bright = rng.random(shape) < bright_frac
    base[bright] += rng.uniform(9000, 40000, bright.sum())
    return np.clip(base, 0, 65535).astype(np.uint16)[..., None]


# ------------------------------------------------------------------ 1. init + forward (C1, C2 nets)
def gold_init_forward():
    for tag, kw in (("c1", dict(coords_channel=3, data_channel=1, layers=5, w0=20, features=22)),
                    ("c2", dict(coords_channel=3, data_channel=1, layers=7, w0=10, features=56)),
                    ("c2small", dict(coords_channel=3, data_channel=1, layers=7, w0=10, features=13)),
                    ("img2d", dict(coords_channel=2, data_channel=1, layers=5, w0=30, features=32))):
        full = dict(kw, name="SIREN", output_act=False, res=False)
        torch.manual_seed(42)
        phi = ref.Networks.init_phi(full)
        nxt = torch.randint(0, 262144, (5,)).numpy()
        torch.manual_seed(42)
        mine = O.init_phi(full)
        nxt2 = torch.randint(0, 262144, (5,)).numpy()
        p_ref, p_me = params_np(phi), params_np(mine)
        for k in p_ref:
            same(p_me[k], p_ref[k], f"init {tag} {k}")
        same(nxt2, nxt, "rng stream after init")
        assert ref.Networks.get_nnmodule_param_count(phi) == O.calc_param_count(**full) == \
            ref.Networks.SIREN.calc_param_count(**full)
        g = torch.Generator().manual_seed(7)
        coords = torch.rand(257, kw["coords_channel"], generator=g) * 2 - 1
        coords[0] = 0
        coords[1] = -1
        coords[2] = 1
        with torch.no_grad():
            y_ref = phi(coords).numpy()
            y_me, zs, acts = O.forward_layers(O.siren_params(mine), coords, kw["w0"])
        same(y_me.numpy(), y_ref, f"forward {tag}")
        sha = hashlib.sha256(b"".join(p.detach().numpy().tobytes() for p in phi.parameters())).hexdigest()
        np.savez_compressed(os.path.join(GOLD, f"siren_{tag}.npz"), coords=coords.numpy(), y=y_ref,
                            next_randint=nxt, sha256=np.array(sha),
                            **{f"z{l}": z.numpy() for l, z in enumerate(zs)},
                            layers=kw["layers"], w0=kw["w0"], features=kw["features"],
                            coords_channel=kw["coords_channel"], **p_ref)
        print("init/forward", tag, sha[:16])


# ------------------------------------------------------------------ 2. width solver
def gold_features():
    rows = []
    for layers in (3, 5, 7, 9):
        for budget in (100.0, 1629.3, 6553.6, 16384.0, 65536.0, 262144.0, 1048576.25):
            kw = dict(coords_channel=3, data_channel=1, layers=layers, res=False)
            f_ref = ref.Networks.SIREN.calc_features(param_count=budget / 4.0, **kw)
            p_ref = ref.Networks.SIREN.calc_param_count(features=f_ref, **kw)
            f_me, p_me, _ = O.estimate_module_size(budget, dict(kw, name="SIREN"))
            assert (f_ref, p_ref) == (f_me, p_me), (layers, budget, f_ref, f_me)
            rows.append((layers, budget, f_ref, p_ref))
    np.savez_compressed(os.path.join(GOLD, "features.npz"), rows=np.array(rows, dtype=np.float64))
    print("features", len(rows))


# ------------------------------------------------------------------ 3. coordinates
def gold_coords():
    out = {}
    for n in (1, 2, 3, 5, 7, 16, 33, 64, 100, 256, 511, 1024):
        for mode in ("-1,1", "n11", "0p1"):
            c_ref = ref.dataset.create_flattened_coords((n, 2), mode)[::2, 0].numpy()
            same(O.axis_coords(n, mode).numpy(), c_ref, f"axis {n} {mode}")
            out[f"axis_{n}_{mode}"] = c_ref
    for shp in ((3, 4, 5), (1, 6, 2), (4, 7)):
        c_ref = ref.dataset.create_flattened_coords(shp, "-1,1").numpy()
        same(O.create_flattened_coords(shp, "-1,1").numpy(), c_ref, f"flat coords {shp}")
        out["flat_" + "x".join(map(str, shp))] = c_ref
    same(O.create_coords((3, 4, 5), "-1,1").numpy(), ref.dataset.create_coords((3, 4, 5), "-1,1").numpy(), "coords")
    np.savez_compressed(os.path.join(GOLD, "coords.npz"), **out)
    print("coords", len(out))


# ------------------------------------------------------------------ 4. normalise / inverse / weights
def gold_normalize():
    blk = synth_block((6, 10, 12), 3)
    t_ref, s_ref = ref.io.normalize_data(blk.copy(), "minmaxany_0_100")
    t_me, s_me = O.normalize_data(blk.copy(), "minmaxany_0_100")
    same(t_me.numpy(), t_ref.numpy(), "normalize")
    assert s_ref == s_me
    probe = torch.tensor([0, 49.9999, 50, 99.9999, 100, 120, -3, 12.3456, 77.7], dtype=torch.float32)
    side = {"dtype": "uint16", "min": 16633.0, "max": 24070.0}
    inv_ref = ref.io.invnormalize_data(probe.clone(), side, "minmaxany_0_100")
    same(O.invnormalize_data(probe.clone(), side, "minmaxany_0_100"), inv_ref, "invnormalize")
    g = torch.Generator().manual_seed(5)
    yhat = torch.rand(4096, generator=g) * 130 - 15
    side2 = {"dtype": "uint16", "min": float(blk.min()), "max": float(blk.max())}
    inv2 = ref.io.invnormalize_data(yhat.clone(), side2, "minmaxany_0_100")
    same(O.invnormalize_data(yhat.clone(), side2, "minmaxany_0_100"), inv2, "invnormalize rand")
    side8 = {"dtype": "uint8", "min": 3.0, "max": 250.0}
    inv8 = ref.io.invnormalize_data(yhat.clone(), side8, "minmaxany_0_100")
    same(O.invnormalize_data(yhat.clone(), side8, "minmaxany_0_100"), inv8, "invnormalize u8")
    w_out = {}
    for i, rule in enumerate((["value_65535_65535_1"], ["value_10001_65535_0.1"],
                              ["value_0_2000_0.5", "value_10001_65535_0.1"], ["none"],
                              ["quantile_1000_0.2_0.9_0.3"])):
        w_ref = ref.misc.parse_weight(blk.copy(), rule)
        same(O.parse_weight(blk.copy(), rule), w_ref, f"parse_weight {rule}")
        w_out[f"w{i}"] = w_ref
    thr_ref, _ = ref.io.normalize_data(np.array(65535), "minmaxany_0_100", max=24070.0, min=16633.0)
    assert float(thr_ref) == O.weight_thres_normalized(65535, "minmaxany_0_100", 16633.0, 24070.0)
    for cp in ("none", "every_2000", "every_7000", "100,300,90000"):
        assert ref.misc.parse_checkpoints(cp, 20000) == O.parse_checkpoints(cp, 20000)
    np.savez_compressed(os.path.join(GOLD, "normalize.npz"), block=blk, normalized=t_ref.numpy(),
                        vmin=s_ref["min"], vmax=s_ref["max"], probe=probe.numpy(), probe_inv=inv_ref,
                        yhat=yhat.numpy(), yhat_inv_u16=inv2, yhat_inv_u8=inv8,
                        thres_norm=float(thr_ref), **w_out)
    print("normalize ok, thres", float(thr_ref))


# ------------------------------------------------------------------ 5. loss + grads + optimiser steps
def gold_train_small():
    """Per-layer gradients and 3 optimiser steps for each optimiser on a replayed index stream."""
    blk = synth_block((8, 12, 10), 11)
    norm = "minmaxany_0_100"
    out = {"block": blk}
    for tag, rule, thres_raw, kw in (
            ("l5", ["value_10001_65535_0.1"], 20000, dict(layers=5, w0=20, features=22)),
            ("l7", ["value_65535_65535_1"], 65535, dict(layers=7, w0=10, features=24))):
        full = dict(coords_channel=3, data_channel=1, name="SIREN", output_act=False, res=False, **kw)
        weight = ref.misc.parse_weight(blk.copy(), rule)
        data_t, side = ref.io.normalize_data(blk.copy(), norm)
        thr, _ = ref.io.normalize_data(np.array(thres_raw), norm, max=side["max"], min=side["min"])
        thr = float(thr)
        for optname in ("Adamax", "Adam", "SGD"):
            torch.manual_seed(42)
            phi = ref.Networks.init_phi(full)
            torch.manual_seed(42)
            mine = O.init_phi(full)
            opt_r = ref.misc.configure_optimizer(phi.parameters(), optname, 1e-3)
            sch_r = ref.misc.configure_lr_scheduler(opt_r, {"name": "MultiStepLR", "milestones": [2, 3], "gamma": 0.2})
            opt_m = O.configure_optimizer(mine.parameters(), optname, 1e-3)
            sch_m = O.configure_lr_scheduler(opt_m, {"name": "MultiStepLR", "milestones": [2, 3], "gamma": 0.2})
            samp_r = O.RandompointSampler(data_t, weight, "-1,1", 300, 4)
            coords_all = ref.dataset.create_flattened_coords(blk.shape[:3], "-1,1")
            idxs, losses = [], []
            out[f"{tag}_{optname}_p0"] = np.concatenate([p.detach().numpy().ravel() for p in phi.parameters()])
            for step, (c, d, w) in enumerate(samp_r):
                idx = samp_r.last_idx
                idxs.append(idx.numpy())
                # reference-side gathers, restated from main.py:154-160 on the reference's coords
                c_ref = coords_all[idx, :]
                same(c.numpy(), c_ref.numpy(), "sampler coords")
                # reference loop body main.py:385-400
                opt_r.zero_grad()
                hat = phi.forward(c_ref)
                w_r = w.clone()
                loss = torch.nn.functional.mse_loss(hat, d, reduction="none")
                if thr:
                    w_r[hat <= thr] = 1
                loss = (loss * w_r).mean()
                loss.backward()
                if step == 0 and optname == "Adamax":
                    gl, gy, gg, gz = O.loss_and_grads(O.siren_params(mine), c, d, w.clone(), thr, kw["w0"])
                    same(gl.numpy(), loss.detach().numpy(), "loss")
                    for l in range(kw["layers"]):
                        same(gg[l][0].numpy(), phi.net[l][0].weight.grad.numpy(), f"dW{l}")
                        same(gg[l][1].numpy(), phi.net[l][0].bias.grad.numpy(), f"db{l}")
                        out[f"{tag}_dW{l}"] = gg[l][0].numpy()
                        out[f"{tag}_db{l}"] = gg[l][1].numpy()
                    out[f"{tag}_yhat0"] = gy.numpy()
                opt_r.step()
                sch_r.step()
                l_me = O.train_step(mine, opt_m, sch_m, c, d, w.clone(), thr)
                same(l_me.detach().numpy(), loss.detach().numpy(), f"{optname} loss step {step}")
                losses.append(float(loss))
                pr = np.concatenate([p.detach().numpy().ravel() for p in phi.parameters()])
                pm = np.concatenate([p.detach().numpy().ravel() for p in mine.parameters()])
                same(pm, pr, f"{optname} params step {step}")
                out[f"{tag}_{optname}_p{step + 1}"] = pr
            out[f"{tag}_{optname}_losses"] = np.array(losses, dtype=np.float64)
            out[f"{tag}_idx"] = np.stack(idxs)
        out[f"{tag}_thr"] = thr
        out[f"{tag}_weight"] = weight
        out[f"{tag}_cfg"] = np.array([kw["layers"], kw["w0"], kw["features"]], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "train_small.npz"), **out)
    print("train_small ok")


# ------------------------------------------------------------------ 6. config 1: 200 full-batch Adamax steps
def gold_config1(steps=200):
    vol, filesize = read_brain()
    np.savez_compressed(os.path.join(GOLD, "brain64.npz"), volume=vol, filesize=filesize)
    norm = "minmaxany_0_100"
    weight = ref.misc.parse_weight(vol.copy(), ["value_65535_65535_1"])
    data_t, side = ref.io.normalize_data(vol.copy(), norm)
    thr, _ = ref.io.normalize_data(np.array(65535), norm, max=side["max"], min=side["min"])
    thr = float(thr)
    phi_kw = dict(coords_channel=3, data_channel=1, layers=5, name="SIREN", w0=20, output_act=False, res=False)
    f, p, theory = O.estimate_module_size(filesize / 80, phi_kw)
    assert (f, p) == (22, 1629), (f, p)
    phi_kw["features"] = f
    torch.manual_seed(42)
    np.random.seed(42)
    phi = ref.Networks.init_phi(phi_kw)
    p0 = np.concatenate([q.detach().numpy().ravel() for q in phi.parameters()])
    opt = ref.misc.configure_optimizer(phi.parameters(), "Adamax", 1e-3)
    sch = ref.misc.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": [50000, 60000, 70000], "gamma": 0.2})
    sampler = O.RandomCubeSampler(data_t, weight, "-1,1", 1, [10000000] * 3, steps)
    assert sampler.pop_size == 1
    losses = []
    for c, d, w in sampler:
        losses.append(float(O.train_step(phi, opt, sch, c, d, w, thr)))
    print("config1 losses", losses[0], losses[49], losses[99], losses[149], losses[-1])
    side_full = dict(side, data_shape=list(data_t.shape), phi_features=f, phi_name="SIREN")
    with tempfile.TemporaryDirectory() as td:
        mp = os.path.join(td, "module")
        ref.ModelSave.save_model(phi, mp, "cpu")
        files = sorted(os.listdir(mp))
        nbytes = sum(os.path.getsize(os.path.join(mp, x)) for x in files)
        mp2 = os.path.join(td, "module2")
        O.save_model(phi, mp2)
        for x in files:
            assert open(os.path.join(mp, x), "rb").read() == open(os.path.join(mp2, x), "rb").read(), x
        phi2 = ref.Networks.init_phi(phi_kw)
        ref.ModelSave.load_model(phi2, mp, "cpu")
        phi3 = O.init_phi(phi_kw)
        O.load_model(phi3, mp2)
        for a, b in zip(phi2.parameters(), phi3.parameters()):
            same(a.detach().numpy(), b.detach().numpy(), "load_model")
    with torch.no_grad():
        rec = ref.misc.reconstruct_flattened(side_full["data_shape"], 10000, phi.forward, device="cpu",
                                             coords_mode="-1,1").float().cpu()
    same(O.reconstruct_flattened(side_full["data_shape"], 10000, phi.forward).numpy(), rec.numpy(), "reconstruct")
    dec = ref.io.invnormalize_data(rec.clone(), side_full, norm)
    same(O.decompress_block(phi, side_full, norm), dec, "decompress")
    psnr = ref.misc.cal_psnr(vol.astype(np.float32), dec.astype(np.float32), 65535)
    ssim = ref.misc.cal_ssim(vol.astype(np.float32), dec.astype(np.float32), 65535)
    assert abs(psnr - O.cal_psnr(vol.astype(np.float32), dec.astype(np.float32), 65535)) < 1e-9
    assert abs(ssim - O.cal_ssim(vol.astype(np.float32), dec.astype(np.float32), 65535)) < 1e-6, \
        (ssim, O.cal_ssim(vol.astype(np.float32), dec.astype(np.float32), 65535))
    print("config1 psnr/ssim", psnr, ssim, "module files", len(files), nbytes)
    pfin = np.concatenate([q.detach().numpy().ravel() for q in phi.parameters()])
    np.savez_compressed(os.path.join(GOLD, "config1_200.npz"), losses=np.array(losses), p0=p0, p_final=pfin,
                        thr=thr, vmin=side["min"], vmax=side["max"], features=f, module_files=np.array(files),
                        module_bytes=nbytes, decompressed=dec, psnr=psnr, ssim=ssim, rec_fp32=rec.numpy())


# ------------------------------------------------------------------ 7. partition helpers
def gold_partition():
    rows = []
    for (d, h, w, nb, ps) in ((64, 512, 512, 4, 262144.0), (64, 512, 512, 64, 262144.0),
                              (1024, 1024, 1024, 64, 4194304.0), (1024, 1024, 1024, -1, 4194304.0),
                              (2048, 2048, 2048, 512, 134217728.0), (60, 90, 75, 12, 5000.0),
                              (64, 64, 64, 1, 100.0)):
        n_ref = ref.adaptive_blocking.cal_divide_num(d, h, w, nb, ps)
        n_me = O.cal_divide_num(d, h, w, nb, ps)
        assert list(n_ref) == list(n_me), (d, h, w, nb, n_ref, n_me)
        rows.append([d, h, w, nb, ps, *n_ref])
    vol = synth_block((12, 20, 18), 21)
    out = {"volume": vol, "divnum": np.array(rows, dtype=np.float64)}
    for dt in ("total_2_2_3", "every_5_8_7"):
        c_ref, _ = ref.misc.divide_data(vol.copy(), dt)
        c_me = O.divide_data(vol.copy(), dt)
        assert [c["name"] for c in c_ref] == [c["name"] for c in c_me]
        for a, b in zip(c_ref, c_me):
            same(b["data"], a["data"], "chunk data")
        for alloc in ("equal", "by_size", "by_var"):
            a_ref = ref.misc.alloc_param([dict(c) for c in c_ref], 9000.0, alloc, 26)
            a_me = O.alloc_param([dict(c) for c in c_me], 9000.0, alloc, 26)
            assert [c["name"] for c in a_ref] == [c["name"] for c in a_me]
            assert [float(c["param_size"]) for c in a_ref] == [float(c["param_size"]) for c in a_me]
            out[f"{dt}_{alloc}_sizes"] = np.array([float(c["param_size"]) for c in a_ref])
        out[f"{dt}_names"] = np.array([c["name"] for c in c_ref])
        m_ref = ref.misc.merge_divided_data(c_ref, vol.shape)
        same(O.merge_divided_data(c_me, vol.shape), m_ref, "merge")
        same(m_ref, vol, "merge roundtrip")
    np.savez_compressed(os.path.join(GOLD, "partition.npz"), **out)
    print("partition ok")


if __name__ == "__main__":
    gold_init_forward()
    gold_features()
    gold_coords()
    gold_normalize()
    gold_train_small()
    gold_partition()
    gold_config1()
    print("all golden fixtures written to", GOLD)
