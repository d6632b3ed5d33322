"""Per-kernel DTW timings (CUDA events through the C ABI's profiler); development tool, not the bench."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import golfer_b200  # noqa: E402

N = int(os.environ.get("QN", "4096"))
T = int(os.environ.get("QT", "300"))
a = torch.randn(N, T, 17, 2, device="cuda").cumsum(1) * 0.05
b = torch.randn(N, T, 17, 2, device="cuda").cumsum(1) * 0.05
ctx = golfer_b200.host._align_ctx(0)
for _ in range(3):
    golfer_b200.host.align_batch(a, b, ctx=ctx)
torch.cuda.synchronize()
ctx.profile(True)
ctx.profile_reset()
for _ in range(5):
    golfer_b200.host.align_batch(a, b, ctx=ctx)
torch.cuda.synchronize()
for k, v in ctx.profile_read().items():
    if v.get("launches"):
        ms = v["ms"] / v["launches"]
        print(f"{k}: {v['launches']} launches, {ms:.3f} ms each" + (f" -> {N / ms * 1e3:.0f} pairs/s" if "dtw" in k else ""))
