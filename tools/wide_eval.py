import os, sys
sys.path.insert(0, "/root/repo")
import torch
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi
for f, prec in ((113, "auto"), (113, "fp32"), (90, "auto")):
    L = 7; dims = (64, 256, 256); nets = 2
    grp = SirenGroup([NetSpec(f, L, 10.0, dims) for _ in range(nets)], 0, prec)
    for j in range(nets):
        torch.manual_seed(42 + j)
        grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
        grp.set_denorm(j, 0.0, 30000.0)
    outs = grp.decompress("uint16"); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): grp.decompress("uint16", out=outs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3; vox = nets * 64 * 256 * 256
    print(f"f={f} {prec}: {ms:.2f} ms  {vox / ms / 1e6:.2f} Gvox/s (SFU bound {4.59e12 / (6 * f) / 1e9:.1f})")
