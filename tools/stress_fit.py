"""Shape sweep of the tensor-core fit / decode kernels against the fp32 kernels on the same inputs (losses, gradients,
decoded values), for deadlock / race hunting: python tools/stress_fit.py [n_cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from brief_pytorch_b200 import _cabi
from brief_pytorch_b200.group import NetSpec, SirenGroup, pack_module_params
from brief_pytorch_b200.Networks import init_phi

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
for case in range(n_cases):
    L = int(rng.choice([3, 4, 5, 7, 8]))
    nets = int(rng.integers(1, 6))
    fs = [int(rng.choice([3, 5, 13, 14, 22, 30, 31, 39, 46, 47, 56, 62, 63, 70, 78, 94, 113, 126])) for _ in range(nets)]
    dims = [tuple(int(x) for x in rng.integers(3, 40, size=3)) for _ in range(nets)]
    modes = [str(rng.choice(["randomcube", "randompoint"])) for _ in range(nets)]
    batches = [int(rng.integers(1, 5000)) for _ in range(nets)]
    steps = int(rng.integers(1, 4))
    res = {}
    for prec in ("fp32", "f16"):
        try:
            grp = SirenGroup([NetSpec(f, L, 10.0, d) for f, d in zip(fs, dims)], 0, prec)
        except _cabi.BriefError as e:
            res = None
            break
        vols = []
        for j in range(nets):
            torch.manual_seed(100 + j)
            grp.set_params(j, pack_module_params(init_phi(dict(name="SIREN", layers=L, w0=10, features=fs[j]))))
            g = torch.Generator().manual_seed(7 + j)
            v = torch.randint(100, 30000, dims[j], dtype=torch.int16, generator=g).cuda()
            vols.append(v)
            grp.bind_volume(j, v, 100.0, 29999.0, np_dtype="uint16", rules=[(10001, 65535, 0.1)], tau=55.0)
            grp.set_sampler(j, modes[j], batches[j])
        losses, grads = [], None
        for s in range(steps):
            losses.append(grp.fit_step(None, seed=5, step=s).cpu().numpy())
            if s == 0:  # gradients are compared on IDENTICAL parameters only: one Adamax step moves every parameter by
                grads = [grp.get_grads(j) for j in range(nets)]  # +-lr, and a near-zero gradient may differ in sign
            if s + 1 < steps:
                grp.opt_step("Adamax", 1e-3)
        dec = [t.cpu().numpy() for t in grp.decompress("float32")]
        res[prec] = (np.stack(losses), grads, dec)
        grp.close()
    if res is None:
        print(f"case {case}: L={L} f={fs} outside the tensor-core envelope, skipped")
        continue
    l32, g32, d32 = res["fp32"]
    l16, g16, d16 = res["f16"]
    el = np.abs(l16 - l32).max() / max(np.abs(l32).max(), 1e-30)
    eg = max(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) for a, b in zip(g16, g32))
    ed = max(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) for a, b in zip(d16, d32))
    worst = max(worst, el, eg, ed)
    # loss and step-0 gradients at the f16 tolerances of the parity tests; the decode after 1-3 optimiser steps sits on
    # two slightly different parameter sets (see above) and is only reported
    flag = "" if el < 1e-2 and eg < 5e-2 and np.isfinite(l16).all() and np.isfinite(ed) else "   <<<<<< CHECK"
    print(f"case {case}: L={L} f={fs} dims={dims} modes={[m[6:] for m in modes]} batch={batches} steps={steps}: "
          f"loss {el:.1e} grad(step 0) {eg:.1e} decode {ed:.1e}{flag}", flush=True)
print("worst relative deviation f16 vs fp32 kernels:", worst)
