"""Per-call wall times of the host-buffer entry point (what bench.py's e2e times); development tool."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import golfer_b200  # noqa: E402

B, T = 256, 300
seg = golfer_b200.Segmenter(golfer_b200.V0, precision="bf16", max_B=B, max_T=T)
x = torch.randn(B, T, 17, 3).pin_memory()
xd = x.cuda()
for _ in range(3):
    seg.segment(xd)
torch.cuda.synchronize()
for mode in ("host", "host", "host_out", "host_out"):
    ts = []
    out = torch.empty((B, T, 9), dtype=torch.float32, pin_memory=True) if mode == "host_out" else None
    for i in range(24):
        t0 = time.perf_counter()
        if out is None:
            r = seg.segment(x)
        else:
            r = seg.segment(x, out=out)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(mode, " ".join(f"{t:.2f}" for t in ts), f"| median {sorted(ts)[12]:.2f} ms -> {B / sorted(ts)[12] * 1e3:.0f} clips/s, mean {sum(ts)/len(ts):.2f}")

# with the bench's NVML sampling thread running
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for period in (0.05, 0.2):
    smp = bench.ClockSampler(0, period)
    smp.start()
    time.sleep(0.3)
    ts = []
    for i in range(40):
        t0 = time.perf_counter()
        r = seg.segment(x)
        ts.append((time.perf_counter() - t0) * 1e3)
    info = smp.stop()
    print(f"nvml {period}s", " ".join(f"{t:.1f}" for t in ts), f"| median {sorted(ts)[20]:.2f} mean {sum(ts)/len(ts):.2f} max {max(ts):.1f}", info)
