"""Per-stage cycle accounting of tc_eval_kernel (CTA 0, epilogue warp 0 and the MMA warp); needs a library built with
-DBRIEF_TC_TIMING:  python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/tc_eval_timing.py 56 7"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brief_pytorch_b200 import _cabi
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi
f = int(sys.argv[1]); L = int(sys.argv[2])
dims = (64, 256, 256)
grp = SirenGroup([NetSpec(f, L, 10.0, dims)], 0, "f16")
torch.manual_seed(42)
grp.set_params(0, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
grp.set_denorm(0, 0.0, 30000.0)
outs = grp.decompress("uint16")
torch.cuda.synchronize()
l = _cabi.load()
buf = (ctypes.c_ulonglong * 64)()
l.brief_debug_read_timing.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
l.brief_debug_read_timing(buf, 1)
grp.decompress("uint16", out=outs)
torch.cuda.synchronize()
l.brief_debug_read_timing(buf, 0)
v = list(buf)
tiles = max(1, v[8]); NH = L - 2
print(f"f={f} L={L} CH={os.environ.get('BRIEF_EVAL_CH', 'default')}: epilogue warp 0 of CTA 0, {tiles} tiles; cycles per tile (per layer step):")
for i, nme in enumerate(["layer 0", "signal (layer 0)", "wait MMA", "LDTM + wait::ld", "bias+sin+pack+STS", "signal", "tile end (y, store)", "TOTAL"]):
    per = NH if 2 <= i <= 5 else 1
    print(f"   {nme:22s} {v[i] / tiles:9.1f}" + (f"   ({v[i] / tiles / per:7.1f} per step)" if per > 1 else ""))
steps = max(1, v[16 + 7])
print(f"MMA warp: per layer step: wait operands {v[16] / steps:7.1f}, issue+commit {v[17] / steps:7.1f}")
