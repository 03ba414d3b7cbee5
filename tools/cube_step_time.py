"""Step time of the general cube sampler against the point sampler on the bench geometry (4 blocks 64x256x256, L=7 f=56,
100 000 samples per block and step): 100 windows of 10x10x10 voxels (index buffer generated per step, fit kernels in
replayed-index mode) vs 100 000 random points drawn on chip.  Random uint16 volumes made on the device.
    python tools/cube_step_time.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brief_pytorch_b200 import Networks  # noqa: E402
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params  # noqa: E402

dims, n, f, L = (64, 256, 256), 4, 56, 7
torch.manual_seed(42)
p0 = pack_module_params(Networks.init_phi(dict(name="SIREN", layers=L, w0=10, features=f)))
raws = [torch.randint(0, 30000, dims, dtype=torch.int16, device="cuda") for _ in range(n)]
out = {}
for mode in ("points", "cubes"):
    grp = SirenGroup([NetSpec(f, L, 10.0, dims) for _ in range(n)], 0, "f16")
    for j in range(n):
        grp.set_params(j, p0)
        grp.bind_volume(j, raws[j], 0.0, 29999.0, 0.0, 100.0, np_dtype="uint16")
        if mode == "cubes":
            grp.set_cube_sampler(j, 100, [10, 10, 10])
        else:
            grp.set_sampler(j, "randompoint", 100000)
    grp.fit_run(30, "Adamax", 1e-3, seed=42)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    grp.fit_run(300, "Adamax", 1e-3, seed=42)
    b.record()
    torch.cuda.synchronize()
    out[mode] = {"ms_per_step": a.elapsed_time(b) / 300, "samples_per_step": n * grp.batch(0)}
    out[mode]["samples_per_s"] = out[mode]["samples_per_step"] / (out[mode]["ms_per_step"] * 1e-3)
    grp.close()
print(json.dumps(out))
