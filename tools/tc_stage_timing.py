"""Per-role cycle accounting of tc_fit_kernel (CTA 0): first warp of group B (forward), first warp of group A
(backward), the MMA-issue warp.  Needs a library built with -DBRIEF_TC_TIMING:
    python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/tc_stage_timing.py [features] [layers] [batch] [nets]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brief_pytorch_b200 import _cabi
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
f = int(argv[0]) if len(argv) > 0 else 56
L = int(argv[1]) if len(argv) > 1 else 7
batch = int(argv[2]) if len(argv) > 2 else 100000
nets = int(argv[3]) if len(argv) > 3 else 4
grp = SirenGroup([NetSpec(f, L, 10.0, (64, 256, 256)) for _ in range(nets)], 0, "f16")
vol = torch.randint(0, 30000, (64, 256, 256), dtype=torch.int16, device="cuda")
for j in range(nets):
    torch.manual_seed(42)
    grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
    grp.bind_volume(j, vol, 0.0, 30000.0, np_dtype="uint16")
    grp.set_sampler(j, "randompoint", batch)
grp.fit_run(5)
torch.cuda.synchronize()
l = _cabi.load()
buf = (ctypes.c_ulonglong * 64)()
l.brief_debug_read_timing.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
l.brief_debug_read_timing(buf, 1)
n_runs = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if "--flush" in sys.argv else None
e0.record()
if flush is None:
    grp.fit_run(n_runs)
else:  # cold L2 before every step, like bench.py's `value`
    for i in range(n_runs):
        flush.fill_(i)
        grp.fit_run(1)
e1.record()
torch.cuda.synchronize()
print(f"f={f} L={L} batch={batch} nets={nets}: {e0.elapsed_time(e1) / n_runs * 1e3:.1f} us per step (instrumented build)")
l.brief_debug_read_timing(buf, 0)
v = list(buf)
NH = L - 2
def show(title, base, names, tiles):
    print(f"{title}: cycles per tile ({tiles / n_runs:.0f} tiles per launch)")
    for i, nme in enumerate(names):
        if nme:
            print(f"   {nme:34s} {v[base + i] / tiles:9.1f}")
tb = max(1, v[8])
show("group B (forward) warp 0", 0, ["wait f2 (tile k-2 done)", "wait sampler", f"wait MMA x{NH + 1}", f"sine epilogue x{NH + 1}",
                                    "y exchange + loss", "wait f1 (dz buffer free)", "dz_NH + signal", "TOTAL"], tb)
ta = max(1, v[16 + 8])
show("group A (backward) warp 0", 16, ["wait loss done (+ dz buffer)", f"wait MMA x{NH}", f"cos epilogue x{NH}", f"signal x{NH}", "dz_NH stage", "", "", "TOTAL"], ta)
print(f"per launch (warp 0): prologue {v[9] / n_runs:.0f} cyc, wait for the other roles at the end {v[10] / n_runs:.0f}, "
      f"gradient flush {v[11] / n_runs:.0f}, kernel total {v[12] / n_runs:.0f}")
tm = max(1, v[32 + 6])
print(f"MMA warp: cycles per tile: A batches {v[32] / tm:.1f} ({v[33] / tm:.1f} batches), B batches/events {v[34] / tm:.1f} ({v[35] / tm:.1f}), "
      f"idle polls {v[36] / tm:.1f}, TOTAL {v[37] / tm:.1f}")
