"""Quick device timings (CUDA events) used while developing; not the bench."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import golfer_b200  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    what = sys.argv[1:] or ["align", "fp32"]
    if "align" in what:
        N = 4096
        a = torch.randn(N, 300, 17, 2, device="cuda").cumsum(1) * 0.05
        b = torch.randn(N, 300, 17, 2, device="cuda").cumsum(1) * 0.05
        ms = timeit(lambda: golfer_b200.host.align_batch(a, b))
        print(f"align 4096x300x300 with path: {ms:.3f} ms -> {N / ms * 1e3:.0f} pairs/s")
        ms = timeit(lambda: golfer_b200.host.align_batch(a, b, want_path=False))
        print(f"align 4096x300x300 cost only: {ms:.3f} ms -> {N / ms * 1e3:.0f} pairs/s")
    for prec in ("fp32", "bf16"):
        if prec not in what:
            continue
        B = int(os.environ.get("QB", "64"))
        seg = golfer_b200.Segmenter(golfer_b200.V0, precision=prec, max_B=B, max_T=300)
        x = torch.randn(B, 300, 17, 3, device="cuda")
        ms = timeit(lambda: seg.segment(x))
        fl = golfer_b200.V0.flops_per_clip(300) * B
        print(f"segment[{prec}] B={B}: {ms:.3f} ms -> {B / ms * 1e3:.0f} clips/s, {fl / ms / 1e9:.1f} TFLOP/s")
        seg.ctx.close()


if __name__ == "__main__":
    main()
