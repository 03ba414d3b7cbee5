"""Long-horizon parity anchor ON THE BENCHED CONFIG (BASELINE configs[1], opt/DivideTask/vessel.yaml as shipped): the
UNMODIFIED reference (imported from /root/reference through oracle/refshim.py) fits ONE 64x256x256 block of the
synthetic vessel volume with SIREN L=7 f=56 w0=10, RandompointSampler batch 100000 (indices from torch's CPU generator,
main.py:156), Adamax lr 1e-3 and the yaml's MultiStepLR scaled from the 80000-step horizon to `steps`
(milestones at 62.5 % / 75 % / 87.5 %, gamma 0.2 — every milestone is crossed).  Recorded in tests/golden/c2_<steps>.npz:
the loss of every step, the final parameters, PSNR / SSIM of the decoded block, and what a replay needs to line up
(seed, a checksum of the volume, the first indices of the first and last step).  The index stream itself is not stored
(2.4 GB): a test replays it with torch.manual_seed(42) -> init_phi -> torch.randint per step, exactly as here.
About 15 minutes on 8 threads.
    python oracle/gen_golden_c2.py [steps]"""
import os
import sys
import zlib

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
torch.set_num_threads(int(os.environ.get("BRIEF_GOLDEN_THREADS", max(1, os.cpu_count() or 1))))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
SHAPE, SEED, BATCH, F, L, W0 = (64, 256, 256), 42, 100000, 56, 7, 10
milestones = [steps * 5 // 8, steps * 6 // 8, steps * 7 // 8]
vol = synth.vessel(SHAPE, seed=SEED)
norm = "minmaxany_0_100"
weight = ref.misc.parse_weight(vol.copy(), ["value_65535_65535_1"])
data_t, side = ref.io.normalize_data(vol.copy(), norm)
thr = O.weight_thres_normalized(65535, norm, side["min"], side["max"])
phi_kw = dict(coords_channel=3, data_channel=1, layers=L, name="SIREN", w0=W0, output_act=False, res=False, features=F)
torch.manual_seed(SEED)
np.random.seed(SEED)
phi = ref.Networks.init_phi(phi_kw)
p0 = np.concatenate([np.concatenate([m[0].weight.detach().numpy().ravel(), m[0].bias.detach().numpy().ravel()]) for m in phi.net])
opt = ref.misc.configure_optimizer(phi.parameters(), "Adamax", 1e-3)
sch = ref.misc.configure_lr_scheduler(opt, {"name": "MultiStepLR", "milestones": milestones, "gamma": 0.2})
sampler = O.RandompointSampler(data_t, weight, "-1,1", BATCH, steps)
losses, idx_first, idx_last = [], None, None
for i, (c, d, w) in enumerate(sampler):
    if i == 0:
        idx_first = sampler.last_idx[:16].numpy().copy()
    idx_last = sampler.last_idx[:16].numpy().copy()
    losses.append(float(O.train_step(phi, opt, sch, c, d, w, thr)))
    if (i + 1) % 100 == 0:
        print(i + 1, losses[-1], flush=True)
p_final = np.concatenate([np.concatenate([m[0].weight.detach().numpy().ravel(), m[0].bias.detach().numpy().ravel()]) for m in phi.net])
side_full = dict(side, data_shape=list(data_t.shape), phi_features=F, phi_name="SIREN")
with torch.no_grad():
    rec = ref.misc.reconstruct_flattened(side_full["data_shape"], 10000, phi.forward, device="cpu", coords_mode="-1,1").float().cpu()
dec = ref.io.invnormalize_data(rec.clone(), side_full, norm)
psnr = ref.misc.cal_psnr(vol.astype(np.float32), dec.astype(np.float32), 65535)
ssim = ref.misc.cal_ssim(vol.astype(np.float32), dec.astype(np.float32), 65535)
print("psnr", psnr, "ssim", ssim)
np.savez_compressed(os.path.join(GOLD, f"c2_{steps}.npz"), steps=steps, shape=np.array(SHAPE), seed=SEED, batch=BATCH,
                    features=F, layers=L, w0=W0, milestones=np.array(milestones), thr=thr,
                    vol_crc32=zlib.crc32(vol.tobytes()), vmin=side["min"], vmax=side["max"],
                    p0=p0.astype(np.float32), p_final=p_final.astype(np.float32), losses=np.array(losses, dtype=np.float64),
                    idx_first=idx_first, idx_last=idx_last, psnr=psnr, ssim=ssim,
                    dec_crc32=zlib.crc32(np.ascontiguousarray(dec).tobytes()))
