"""Quality anchor for the WIDE kernels: the UNMODIFIED reference (imported from /root/reference through
oracle/refshim.py) fits a wide SIREN (L = 5, f = 70, w0 = 20 — F_PAD = 80, the wide tcgen05 fit kernel's bucket) on the
shipped 64^3 brain block with SingleTask default.yaml's other settings (full-batch Adamax) for 600 steps on the CPU; the
loss every 50 steps and the final PSNR / SSIM go to tests/golden/wide70_600.npz.  A few minutes on 8 threads.
    python oracle/gen_golden_wide.py [steps]"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import brief_oracle as O  # noqa: E402
import refshim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
ref = refshim.load_reference()
torch.set_num_threads(max(1, os.cpu_count() or 1))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
vol = np.load(os.path.join(GOLD, "brain64.npz"))["volume"]
g200 = np.load(os.path.join(GOLD, "config1_200.npz"))
norm = "minmaxany_0_100"
weight = ref.misc.parse_weight(vol.copy(), ["value_65535_65535_1"])
data_t, side = ref.io.normalize_data(vol.copy(), norm)
thr = float(g200["thr"])
F = 70
phi_kw = dict(coords_channel=3, data_channel=1, layers=5, name="SIREN", w0=20, output_act=False, res=False, features=F)
torch.manual_seed(42)
np.random.seed(42)
phi = ref.Networks.init_phi(phi_kw)
p0 = np.concatenate([np.concatenate([m[0].weight.detach().numpy().ravel(), m[0].bias.detach().numpy().ravel()]) for m in phi.net])
opt = ref.misc.configure_optimizer(phi.parameters(), "Adamax", 1e-3)
sch = ref.misc.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": [50000, 60000, 70000], "gamma": 0.2})
sampler = O.RandomCubeSampler(data_t, weight, "-1,1", 1, [10000000] * 3, steps)
losses = []
for i, (c, d, w) in enumerate(sampler):
    losses.append(float(O.train_step(phi, opt, sch, c, d, w, thr)))
    if (i + 1) % 50 == 0:
        print(i + 1, losses[-1], flush=True)
side_full = dict(side, data_shape=list(data_t.shape), phi_features=F, phi_name="SIREN")
with torch.no_grad():
    rec = ref.misc.reconstruct_flattened(side_full["data_shape"], 10000, phi.forward, device="cpu", coords_mode="-1,1").float().cpu()
dec = ref.io.invnormalize_data(rec.clone(), side_full, norm)
psnr = ref.misc.cal_psnr(vol.astype(np.float32), dec.astype(np.float32), 65535)
ssim = ref.misc.cal_ssim(vol.astype(np.float32), dec.astype(np.float32), 65535)
print("psnr", psnr, "ssim", ssim)
np.savez_compressed(os.path.join(GOLD, f"wide{F}_{steps}.npz"), steps=steps, features=F, p0=p0.astype(np.float32), thr=thr,
                    losses=np.array(losses), psnr=psnr, ssim=ssim)
