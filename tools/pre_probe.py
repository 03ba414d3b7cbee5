"""Time brief_preprocess on one 1024^3 uint16 block (2 GiB).  Plain run prints CUDA-event times; under
`ncu --metrics gpu__time_duration.sum` the launch list gives each pass's share."""
import sys

import torch

sys.path.insert(0, ".")
from brief_pytorch_b200.group import preprocess_  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
close = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 2, 2]
vol = torch.empty((n, 1024, 1024), dtype=torch.int16, device="cuda")
vol.random_(0, 30000)
vol[: n // 2] >>= 6
scratch = torch.empty(2 * n * 1024 * 32 * 4, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for clip in ([0, 65535], [100, 20000]):
    for it in range(4):
        e0.record()
        preprocess_(vol, 500, close, clip, "uint16", scratch=scratch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"close {close} clip {clip}: {ms:.3f} ms  {vol.numel() / ms / 1e6:.1f} Gvox/s  read-once {vol.numel() * 2 / ms / 1e6:.0f} GB/s")
