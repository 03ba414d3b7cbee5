"""Per-stage cycle accounting of tc_fit_wide_kernel (CTA 0, warp 0).  Needs a library built with -DBRIEF_TC_TIMING:
    python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/wide_timing.py [features] [layers] [nets]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brief_pytorch_b200 import _cabi
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi
f = int(sys.argv[1]) if len(sys.argv) > 1 else 113
L = int(sys.argv[2]) if len(sys.argv) > 2 else 7
nets = int(sys.argv[3]) if len(sys.argv) > 3 else 4
grp = SirenGroup([NetSpec(f, L, 10.0, (256, 256, 256)) for _ in range(nets)], 0, "f16")
vol = torch.randint(0, 30000, (256, 256, 256), dtype=torch.int16, device="cuda")
for j in range(nets):
    torch.manual_seed(42)
    grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
    grp.bind_volume(j, vol, 0.0, 30000.0, np_dtype="uint16")
    grp.set_sampler(j, "randompoint", 100000)
grp.fit_run(3)
torch.cuda.synchronize()
l = _cabi.load()
buf = (ctypes.c_ulonglong * 64)()
l.brief_debug_read_timing_wide.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
l.brief_debug_read_timing_wide(buf, 1)
n_runs = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); grp.fit_run(n_runs); e1.record()
torch.cuda.synchronize()
print(f"f={f} L={L} nets={nets}: {e0.elapsed_time(e1) / n_runs * 1e3:.1f} us per step (instrumented build)")
l.brief_debug_read_timing_wide(buf, 0)
v = list(buf)
tiles = max(1, v[12])
names = ["sampler + sync", f"fwd: issue + MMA wait x{L - 1}", f"fwd: sine epilogue x{L - 1}", f"fwd: sync x{L - 1}", "loss + dz_NH + sync",
         "dWlast (MMA, drain, sync)", f"bwd: issue + MMA wait x{L - 2}", f"bwd: cos epilogue x{L - 2}", f"bwd: wait dW x{L - 2}",
         f"bwd: drain x{L - 2}", f"bwd: sync x{L - 2}", "dW0 (MMA, drain, sync)", "", "TOTAL per tile"]
print(f"cycles per tile ({tiles / n_runs:.1f} tiles per launch in CTA 0)")
for i, nme in enumerate(names):
    if nme:
        print(f"   {nme:34s} {v[i] / tiles:9.1f}")
