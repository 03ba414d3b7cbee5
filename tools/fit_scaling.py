"""Fit-kernel time vs samples per step: slope = per-tile cost, intercept = per-launch fixed cost (weights staging,
TMEM alloc, partial flush).  python tools/fit_scaling.py [features] [layers]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi

f = int(sys.argv[1]) if len(sys.argv) > 1 else 56
L = int(sys.argv[2]) if len(sys.argv) > 2 else 7
nets = 4
vol = torch.randint(0, 30000, (64, 256, 256), dtype=torch.int16, device="cuda")
for batch in (148 * 128 // nets * 2, 148 * 128 // nets * 8, 100000, 200000, 400000):
    grp = SirenGroup([NetSpec(f, L, 10.0, (64, 256, 256)) for _ in range(nets)], 0, "f16")
    for j in range(nets):
        torch.manual_seed(42)
        grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
        grp.bind_volume(j, vol, 0.0, 30000.0, np_dtype="uint16")
        grp.set_sampler(j, "randompoint", batch)
    for s in range(5):
        grp.fit_kernel_only(seed=1, step=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(20):
        grp.fit_kernel_only(seed=1, step=s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tiles = nets * ((batch + 127) // 128)
    print(f"f={f} L={L} batch/net={batch:7d} tiles={tiles:6d} tiles/SM={tiles/148:6.1f}  kernel {ms*1e3:8.1f} us  "
          f"{nets*batch/ms/1e6:8.1f} Msamples/s")
    grp.close()
