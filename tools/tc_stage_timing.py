import ctypes, torch, numpy as np, sys, os
sys.path.insert(0, '/root/repo')
os.environ["BRIEF_NO_BUILD"] = "1"
from brief_pytorch_b200 import build as _b
_b.LIBPATH = "/root/repo/brief_pytorch_b200/_lib/libbrief_timing.so"
from brief_pytorch_b200 import _cabi
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi
grp = SirenGroup([NetSpec(56, 7, 10.0, (64,256,256)) for _ in range(4)], 0, "f16")
vol = torch.randint(0, 30000, (64,256,256), dtype=torch.int16, device="cuda")
for j in range(4):
    torch.manual_seed(42)
    grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=7, w0=10, features=56))))
    grp.bind_volume(j, vol, 0.0, 30000.0, np_dtype="uint16")
    grp.set_sampler(j, "randompoint", 100000)
grp.fit_run(5)
torch.cuda.synchronize()
lib = _cabi.load()
buf = (ctypes.c_ulonglong * 16)()
rt = ctypes.CDLL("libcudart.so.12")
# read g_tc_timing via cudaMemcpyFromSymbol is awkward from ctypes; use helper exported by the lib
lib.brief_debug_read_timing.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
lib.brief_debug_read_timing(buf, 1)
grp.fit_run(10)
torch.cuda.synchronize()
lib.brief_debug_read_timing(buf, 0)
v = list(buf)
names = ["fence", "syncthreads", "issue(t0)/skip", "mbar_wait", "tmem_ld", "sin+store", "tile_total", "tiles"]
for who, off in (("t0", 0), ("t511", 8)):
    tiles = v[off + 7]
    print(who, "tiles", tiles)
    for i, nme in enumerate(names[:6]):
        print(f"   fwd stage {nme:16s}: {v[off+i]/ (tiles*5):8.1f} cyc/stage")
    print(f"   tile total: {v[off+6]/tiles:9.1f} cyc")
