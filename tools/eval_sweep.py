"""Decompress-kernel timing against the SFU bound: python tools/eval_sweep.py f L [nets]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi
f = int(sys.argv[1]); L = int(sys.argv[2]); nets = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dims = (64, 256, 256)
grp = SirenGroup([NetSpec(f, L, 10.0, dims) for _ in range(nets)], 0, "f16")
for j in range(nets):
    torch.manual_seed(42 + j)
    grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
    grp.set_denorm(j, 0.0, 30000.0)
outs = grp.decompress("uint16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    grp.decompress("uint16", out=outs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
vox = nets * dims[0] * dims[1] * dims[2]
sfu = 4.59e12 / ((L - 1) * f)
print(f"f={f} L={L}: {ms:.3f} ms  {vox / ms / 1e6:.2f} Gvox/s  ({100 * vox / ms / 1e-3 / sfu:.1f}% of the SFU bound {sfu / 1e9:.1f} Gvox/s)  checksum {int(outs[0].view(torch.int16).to(torch.int64).sum())}")
