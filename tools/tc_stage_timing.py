"""Per-stage cycle accounting of tc_fit_kernel (CTA 0): builds a second library with -DBRIEF_TC_TIMING and prints where
an epilogue warp, the MMA-issue warp and the sampler warp spend their cycles.
    python tools/tc_stage_timing.py [features] [layers] [batch] [nets]      (on a GPU box)"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brief_pytorch_b200 import build as _b
lib = os.path.join(_b.LIBDIR, "libbrief_timing.so")
flags = [f for f in _b.NVCC_FLAGS if f != "--use_fast_math=false"]
if "--no-build" not in sys.argv:
    subprocess.run([_b._nvcc(), *flags, "-DBRIEF_TC_TIMING", "-o", lib] + [os.path.join(_b.CSRC, s) for s in _b.SOURCES], check=True)
if "--build-only" in sys.argv:
    sys.exit(0)
os.environ["BRIEF_NO_BUILD"] = "1"
_b.LIBPATH = lib
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
e0.record(); grp.fit_run(n_runs); e1.record()
torch.cuda.synchronize()
print(f"f={f} L={L} batch={batch} nets={nets}: {e0.elapsed_time(e1) / n_runs * 1e3:.1f} us per step (instrumented build)")
l.brief_debug_read_timing(buf, 0)
v = list(buf)
NH = L - 2
tiles = max(1, v[8])
print(f"epilogue warp 0 (CTA 0): {tiles / n_runs:.0f} tiles per launch; cycles per tile:")
for i, nme in enumerate(["fwd_prologue+signal", "wait A (z,dX)", "bwd_epilogue+signal", "wait B (fwd)", "fwd_epilogue+signal",
                         "wait A (tile end)", "loss_phase", "TOTAL"]):
    per = NH if 1 <= i <= 4 else 1
    print(f"   {nme:22s} {v[i] / tiles:9.1f}   ({v[i] / tiles / per:7.1f} per stage)" if per > 1 else f"   {nme:22s} {v[i] / tiles:9.1f}")
mt = max(1, v[16 + 7])
print("MMA warp: cycles per tile:")
for i, nme in enumerate(["wait ra (tile start)", "issue bwd NH + fwd 1", "wait ra (stage)", "issue bwd stage", "wait rb (stage)",
                         "issue fwd stage", "TOTAL"]):
    print(f"   {nme:22s} {v[16 + i] / mt:9.1f}")
st = max(1, v[32 + 7])
print("sampler warp: cycles per tile:")
for i, nme in enumerate(["wait gfree", "sample_tile"]):
    print(f"   {nme:22s} {v[32 + i] / st:9.1f}")
