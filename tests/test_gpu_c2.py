"""Long-horizon parity ON THE BENCHED CONFIG (BASELINE configs[1], vessel.yaml as shipped: one 64x256x256 block, SIREN L=7
f=56 w0=10, RandompointSampler batch 100000, Adamax, MultiStepLR) against tests/golden/c2_3000.npz, which
oracle/gen_golden_c2.py wrote from the UNMODIFIED reference run on the CPU for 3000 steps (milestones scaled to
62.5 / 75 / 87.5 % of the horizon, so every one is crossed).

 * replayed: the reference's own torch.randint index stream (same seed -> init_phi -> randint order), every step through
   the C-ABI.  Loss curve: within 3 x TOL while the trajectories still coincide, window means within 5 % later;
   decoded PSNR within 0.1 dB and SSIM within 0.002 of the reference's (utils/misc.py:447-499, utils/ssim.py) at equal steps.
 * philox: what bench.py's `value` times — the on-device sampler.  A different random stream, so only the statistics
   can agree: the same PSNR / SSIM bars.
"""
import zlib

import numpy as np
import pytest
import torch

import brief_oracle as O
from conftest import load_gold

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "f16": 1e-2}


def _setup(prec):
    from brief_pytorch_b200 import Networks, synth
    from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
    g = load_gold("c2_3000")
    shape = tuple(int(x) for x in g["shape"])
    vol = synth.vessel(shape, seed=int(g["seed"]))
    assert zlib.crc32(vol.tobytes()) == int(g["vol_crc32"]), "synthetic volume differs from the one the reference was run on"
    kw = dict(coords_channel=3, data_channel=1, layers=int(g["layers"]), name="SIREN", w0=int(g["w0"]), features=int(g["features"]))
    torch.manual_seed(int(g["seed"]))
    phi = Networks.init_phi(kw)  # consumes the CPU generator exactly like the reference's init_phi
    p0 = pack_module_params(phi)
    np.testing.assert_array_equal(p0, g["p0"])
    grp = SirenGroup([NetSpec(kw["features"], kw["layers"], kw["w0"], shape)], 0, prec)
    assert grp.precision(0) == prec
    grp.set_axes(0, "-1,1")
    grp.set_params(0, p0)
    raw = torch.from_numpy(np.ascontiguousarray(vol[..., 0]).view(np.int16)).cuda()
    grp.bind_volume(0, raw, float(g["vmin"]), float(g["vmax"]), 0.0, 100.0, rules=[(65535, 65535, 1.0)], tau=float(g["thr"]),
                    np_dtype="uint16")
    grp.set_sampler(0, "randompoint", int(g["batch"]))
    return g, vol, grp, raw


def _quality(g, vol, grp):
    dec = grp.decompress("uint16")[0].cpu().numpy().view(np.uint16)[..., None]
    a, b = vol.astype(np.float32), dec.astype(np.float32)
    return O.cal_psnr(a, b, 65535), O.cal_ssim(a, b, 65535)


@pytest.mark.parametrize("prec", ["f16", "fp32"])
def test_c2_replayed_index_stream_follows_the_reference_for_3000_steps(prec):
    g, vol, grp, raw = _setup(prec)
    steps, batch, n_vox = int(g["steps"]), int(g["batch"]), vol.size
    ms = [int(m) for m in g["milestones"]]
    # two pinned buffer sets: step s+1's indices are drawn and staged while step s runs (brief_fit_step_host graphs)
    idx = [torch.empty(batch, dtype=torch.int64).pin_memory() for _ in range(2)]
    loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    losses = np.zeros(steps)
    for s in range(steps):
        k = s & 1
        if s >= 2:
            done[k].synchronize()
            losses[s - 2] = float(loss[k][0])
        drawn = torch.randint(0, n_vox, (batch,))  # the reference's stream (main.py:156), same generator state
        if s == 0:
            np.testing.assert_array_equal(drawn[:16].numpy(), g["idx_first"])
        if s == steps - 1:
            np.testing.assert_array_equal(drawn[:16].numpy(), g["idx_last"])
        idx[k].copy_(drawn)
        grp.fit_step_host(idx[k], loss[k], "Adamax", 1e-3, milestones=ms, gamma=0.2)
        done[k].record()
    torch.cuda.synchronize()
    losses[steps - 2], losses[steps - 1] = float(loss[steps & 1][0]), float(loss[(steps - 1) & 1][0])
    ref = g["losses"]
    tol = TOL[prec]
    assert np.isfinite(losses).all()
    # the first steps: same parameters, same samples -> the same loss to the arithmetic's tolerance
    np.testing.assert_allclose(losses[:20], ref[:20], rtol=3 * tol)
    # later the trajectories separate through rounding alone; the window means of the loss still agree
    for lo in range(0, steps, 250):
        a, b = losses[lo:lo + 250].mean(), ref[lo:lo + 250].mean()
        assert abs(a - b) <= 0.05 * b, (lo, a, b)
    psnr, ssim = _quality(g, vol, grp)
    print(f"c2 replay [{prec}]: psnr {psnr:.3f} (reference {float(g['psnr']):.3f}), ssim {ssim:.5f} ({float(g['ssim']):.5f}), "
          f"final loss {losses[-50:].mean():.4f} ({ref[-50:].mean():.4f})")
    assert abs(psnr - float(g["psnr"])) <= 0.1, (psnr, float(g["psnr"]))
    assert abs(ssim - float(g["ssim"])) <= 0.002, (ssim, float(g["ssim"]))
    grp.close()


def test_c2_on_device_sampler_reaches_the_reference_quality():
    """bench.py's timed configuration: Philox sampler on the device, whole run enqueued from C (brief_fit_run)."""
    g, vol, grp, raw = _setup("f16")
    ms = [int(m) for m in g["milestones"]]
    hist = grp.fit_run(int(g["steps"]), "Adamax", 1e-3, milestones=ms, gamma=0.2, seed=42, loss_history=True)
    losses = hist[:, 0].cpu().numpy().astype(np.float64)
    ref = g["losses"]
    for lo in range(250, int(g["steps"]), 250):  # another sample stream: compare window means only, loosely
        a, b = losses[lo:lo + 250].mean(), ref[lo:lo + 250].mean()
        assert abs(a - b) <= 0.10 * b, (lo, a, b)
    psnr, ssim = _quality(g, vol, grp)
    print(f"c2 philox [f16]: psnr {psnr:.3f} (reference {float(g['psnr']):.3f}), ssim {ssim:.5f} ({float(g['ssim']):.5f})")
    assert abs(psnr - float(g["psnr"])) <= 0.1, (psnr, float(g["psnr"]))
    assert abs(ssim - float(g["ssim"])) <= 0.002, (ssim, float(g["ssim"]))
    grp.close()
