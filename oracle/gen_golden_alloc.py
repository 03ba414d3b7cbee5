"""Budget allocation 'by_d' / 'by_dv' (utils/misc.py:410-422 with the 3-D FFT feature of utils/adaptive_blocking.py:16-24)
from the UNMODIFIED reference (imported through oracle/refshim.py) on the partition fixture's small volume and on a
textured one; written to tests/golden/alloc_d.npz.
    python oracle/gen_golden_alloc.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import refshim  # noqa: E402
from brief_pytorch_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ref = refshim.load_reference()
out = {}
vols = {"small": np.load(os.path.join(GOLD, "partition.npz"))["volume"], "hipct": synth.hipct((24, 32, 40), seed=5)}
for tag, vol in vols.items():
    out[f"{tag}_volume"] = vol
    for dt in ("total_2_2_2", "every_7_9_11"):
        chunks, _ = ref.misc.divide_data(vol.copy(), dt)
        out[f"{tag}_{dt}_feature"] = np.array([ref.adaptive_blocking.cal_feature(c["data"]) for c in chunks])
        for alloc in ("by_d", "by_dv"):
            kept = ref.misc.alloc_param([dict(c) for c in chunks], 9000.0, alloc, 26)
            out[f"{tag}_{dt}_{alloc}_names"] = np.array([c["name"] for c in kept])
            out[f"{tag}_{dt}_{alloc}_sizes"] = np.array([float(c["param_size"]) for c in kept])
np.savez_compressed(os.path.join(GOLD, "alloc_d.npz"), **out)
print("alloc_d ok", {k: v.shape for k, v in out.items() if k.endswith("sizes")})
