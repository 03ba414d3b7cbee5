"""Parity anchor for the GENERAL RandomCubeSampler (cube_len smaller than the block, cube_count > 1) and for lr schedules
with more than 8 decays (StepLR): the UNMODIFIED reference — its live sampler class cut out of main.py:38-125
(refshim.load_main_samplers), utils.Networks / utils.misc / utils.io imported through oracle/refshim.py — fits

  c3d: a 20x24x28 uint16 neuron block, SIREN L=5 f=24 w0=20, 5 cubes of 8x6x10 voxels per step (population 4693 windows),
       weight rule value_10001_65535_0.1, weight_thres 40 (normalised), Adamax + MultiStepLR([20, 30], 0.2), 40 steps;
  c2d: a 40x48 uint8 image (coords_channel 2), SIREN L=4 f=16 w0=30, 6 windows of 7x9 pixels per step, Adam +
       StepLR(step_size 3, gamma 0.7) over 30 steps (9 decays: more than one BriefOptConfig window).

Recorded in tests/golden/cubes.npz: the block, every step's window draws (torch.randint on the CPU generator,
main.py:114), the first step's three sampler outputs, the loss of every step, the initial and final parameters and the
learning rate of every step.
    python oracle/gen_golden_cubes.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import brief_oracle as O  # noqa: E402
import refshim  # noqa: E402
from brief_pytorch_b200 import synth  # noqa: E402  (numpy-only generator of the synthetic volume)

GOLD = os.path.join(ROOT, "tests", "golden")
ref = refshim.load_reference()
live = refshim.load_main_samplers()
torch.set_num_threads(1)
NORM = "minmaxany_0_100"


def packed(phi):
    return np.concatenate([np.concatenate([m[0].weight.detach().numpy().ravel(), m[0].bias.detach().numpy().ravel()])
                           for m in phi.net]).astype(np.float32)


def run(tag, vol, phi_kw, rules, tau, cube_count, cube_len, steps, optname, sched, out):
    weight = ref.misc.parse_weight(vol.copy(), rules)
    data_t, side = ref.io.normalize_data(vol.copy(), NORM)
    torch.manual_seed(42)
    np.random.seed(42)
    phi = ref.Networks.init_phi(dict(phi_kw))
    out[f"{tag}_p0"] = packed(phi)
    opt = ref.misc.configure_optimizer(phi.parameters(), optname, 1e-3)
    sch = ref.misc.configure_lr_scheduler(opt, sched)
    sampler = live.RandomCubeSampler(data_t, weight, "-1,1", cube_count, list(cube_len), steps)
    mirror = O.RandomCubeSampler(data_t, weight, "-1,1", cube_count, list(cube_len), steps)
    assert sampler.pop_size == mirror.pop_size
    losses, lrs, ids = [], [], []
    it = iter(sampler)
    iter(mirror)
    for i in range(steps):
        state = torch.get_rng_state()
        c, d, w = next(it)
        torch.set_rng_state(state)       # the oracle's restatement draws the same windows ...
        c2, d2, w2 = next(mirror)
        assert torch.equal(c, c2) and torch.equal(d, d2) and torch.equal(w, w2)   # ... and returns the same tensors
        ids.append(mirror.last_idx.numpy().copy())
        if i == 0:
            out[f"{tag}_coords0"], out[f"{tag}_data0"], out[f"{tag}_weight0"] = c.numpy().copy(), d.numpy().copy(), w.numpy().copy()
        lrs.append(opt.param_groups[0]["lr"])
        losses.append(float(O.train_step(phi, opt, sch, c, d, w, tau)))
    print(tag, "pop", sampler.pop_size, "loss", losses[0], "->", losses[-1], "lr", lrs[0], "->", lrs[-1])
    out.update({f"{tag}_vol": vol, f"{tag}_ids": np.stack(ids), f"{tag}_losses": np.array(losses, dtype=np.float64),
                f"{tag}_lrs": np.array(lrs, dtype=np.float64), f"{tag}_p_final": packed(phi), f"{tag}_vmin": side["min"],
                f"{tag}_vmax": side["max"], f"{tag}_pop": sampler.pop_size, f"{tag}_tau": tau,
                f"{tag}_cube_len": np.array(sampler.data_cubes.shape[1:-1]), f"{tag}_cube_count": cube_count})


out = {}
run("c3d", synth.neuron((20, 24, 28), seed=5),
    dict(coords_channel=3, data_channel=1, layers=5, name="SIREN", w0=20, output_act=False, res=False, features=24),
    ["value_10001_65535_0.1"], 40.0, 5, [8, 6, 10], 40, "Adamax",
    {"name": "MultiStepLR", "milestones": [20, 30], "gamma": 0.2}, out)
rng = np.random.default_rng(8)
yy, xx = np.meshgrid(np.arange(40), np.arange(48), indexing="ij")
img = np.clip(120 + 90 * np.sin(yy / 5.0) * np.cos(xx / 7.0) + rng.integers(-12, 13, size=(40, 48)), 0, 255).astype(np.uint8)[..., None]
run("c2d", img,
    dict(coords_channel=2, data_channel=1, layers=4, name="SIREN", w0=30, output_act=False, res=False, features=16),
    ["none"], 0.0, 6, [7, 9], 30, "Adam", {"name": "StepLR", "step_size": 3, "gamma": 0.7}, out)
np.savez_compressed(os.path.join(GOLD, "cubes.npz"), **out)
print("wrote", os.path.join(GOLD, "cubes.npz"))
