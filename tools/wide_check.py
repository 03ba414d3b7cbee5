"""Wide tensor-core fit kernel (64 < F_PAD <= 128) against the fp32 kernels on the same inputs: loss, every gradient
tensor, then a timing of the two on hipct-shaped work.  python tools/wide_check.py [quick]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi

quick = len(sys.argv) > 1
cases = [(113, 7, (5, 24, 40), "randomcube", 0), (70, 7, (3, 10, 9), "randomcube", 0), (90, 5, (7, 33, 31), "randompoint", 3000),
         (126, 7, (16, 32, 32), "randompoint", 5000), (113, 3, (4, 16, 16), "randomcube", 0), (100, 4, (4, 16, 17), "randomcube", 0),
         ]
for f, L, dims, mode, batch in cases:
    res = {}
    for prec in ("fp32", "f16"):
        grp = SirenGroup([NetSpec(f, L, 10.0, dims), NetSpec(f, L, 10.0, dims)], 0, prec)
        for j in range(2):
            torch.manual_seed(100 + j)
            grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=f))))
            g = torch.Generator().manual_seed(7 + j)
            v = torch.randint(100, 30000, dims, dtype=torch.int16, generator=g).cuda()
            grp.bind_volume(j, v, 100.0, 29999.0, np_dtype="uint16", rules=[(10001, 65535, 0.1)], tau=55.0)
            grp.set_sampler(j, mode, batch or 1)
        assert grp.precision(0) == prec, grp.precision(0)
        loss = grp.fit_step(None, seed=5, step=0).cpu().numpy()
        torch.cuda.synchronize()
        res[prec] = (loss, [grp.get_grads(j) for j in range(2)])
        grp.close()
    l32, g32 = res["fp32"]
    l16, g16 = res["f16"]
    el = np.abs(l16 - l32).max() / np.abs(l32).max()
    eg = max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(g16, g32))
    print(f"f={f} L={L} dims={dims} {mode} batch={batch}: loss {l32} vs {l16} ({el:.1e}), grad rel dev {eg:.1e}"
          f"{'' if el < 1e-2 and eg < 5e-2 else '   <<<<<< FAIL'}", flush=True)

if not quick:
    for f, nets in ((113, 4), (113, 16), (78, 8)):
        for prec in ("fp32", "f16"):
            dims = (256, 256, 256)
            grp = SirenGroup([NetSpec(f, 7, 10.0, dims) for _ in range(nets)], 0, prec)
            vols = []
            for j in range(nets):
                torch.manual_seed(100 + j)
                grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=7, w0=10, features=f))))
                v = torch.randint(100, 30000, dims, dtype=torch.int16, device="cuda")
                vols.append(v)
                grp.bind_volume(j, v, 100.0, 29999.0, np_dtype="uint16", rules=[(65535, 65535, 1.0)], tau=0.0)
                grp.set_sampler(j, "randompoint", 100000)
            for s in range(3):
                grp.fit_step(None, seed=5, step=s); grp.opt_step("Adamax", 1e-3)
            torch.cuda.synchronize()
            n = 10 if prec == "f16" else 3
            t0 = time.perf_counter()
            for s in range(n):
                loss = grp.fit_step(None, seed=5, step=3 + s); grp.opt_step("Adamax", 1e-3)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            print(f"f={f} nets={nets} {prec}: {dt * 1e3:.3f} ms/step  {nets * 100000 / dt / 1e6:.1f} M samples/s  loss {float(loss[0]):.3f}", flush=True)
            grp.close()
