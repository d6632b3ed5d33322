"""Dump the tc_gemm kernel's clock64 trace (CTA 0, first tiles) — GOLFER_TRACE_TC=1.

roles: 0 producer (ev c: chunk c issued), 1 mma (ev c: chunk c issued, 31: accumulator committed),
2 residual producer (ev q: box q handed over), 3/4 epilogue group 0/1 (0: accumulator ready;
per box q: 1+4q slot ready, 2+4q TMEM loaded, 3+4q staged + barrier, 4+4q pooling sums done)."""
import os
import sys

os.environ["GOLFER_TRACE_TC"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import golfer_b200  # noqa: E402

B = int(os.environ.get("QB", "256"))
blocks = [int(x) for x in os.environ.get("BLOCKS", "1,5").split(",")]
seg = golfer_b200.Segmenter(golfer_b200.V0, precision="bf16", max_B=B, max_T=300)
x = torch.randn(B, 300, 17, 3, device="cuda")
for _ in range(2):
    seg.segment(x)
torch.cuda.synchronize()
raw = seg.ctx.debug_read("tc_trace", 8 * 2 * 5 * 8 * 32 * 8).view(np.uint64).reshape(8, 2, 5, 8, 32).astype(np.int64)
for blk in blocks:
    for kern, kname in enumerate(("tcn1x1", "tconv")):
        t = raw[blk, kern]
        if not (t > 0).any():
            continue
        t0 = t[t > 0].min()
        C = golfer_b200.V0.widths[blk]
        nq = C // 64
        print(f"=== block {blk} {kname} (C={C}); cycles relative to first event")
        rel = lambda a: [int(v - t0) if v > 0 else -1 for v in a]
        if kname == "tconv":
            # tconv_window.cuh roles: 0 producer [v][g], 1 mma [n][g], 3+g epilogue [n]: 5 wait acc, 6 acc ready,
            # 0 TMEM loaded, 1 staged, 2 barrier passed, 3 pooling done
            for n in range(8):
                print(f" step {n}: mma g0: top {rel(t[1, n, 6:7])} acc-free {rel(t[1, n, 2:3])} data {rel(t[1, n, 4:5])} issued {rel(t[1, n, 0:1])}"
                      f" | g1: top {rel(t[1, n, 7:8])} acc-free {rel(t[1, n, 3:4])} data {rel(t[1, n, 5:6])} issued {rel(t[1, n, 1:2])}")
            for g in (0, 1):
                for n in range(8):
                    e = t[3 + g, n]
                    print(f"   group {g} step {n}: wait {rel(e[5:6])} acc {rel(e[6:7])} ld {rel(e[0:1])} staged {rel(e[1:2])} "
                          f"bar {rel(e[2:3])} pooled {rel(e[3:4])}")
            continue
        for tile in range(1, 5):
            nz = int((t[0, tile] > 0).sum())
            print(f" tile {tile}: prod {rel(t[0, tile, :nz])}")
            print(f"          mma  {rel(t[1, tile, :nz])} commit {rel(t[1, tile, 31:32])}")
            print(f"          res  {rel(t[2, tile, :nq])}")
        for g in (0, 1):
            for k in range(0, 3):
                print(f" group {g} tile#{k}: acc ready {rel(t[3 + g, k, 0:1])}")
                for q in range(nq):
                    print(f"     box {q}: slot {rel(t[3 + g, k, 1 + 4 * q:2 + 4 * q])} ld {rel(t[3 + g, k, 2 + 4 * q:3 + 4 * q])} "
                          f"staged {rel(t[3 + g, k, 3 + 4 * q:4 + 4 * q])} stored {rel(t[3 + g, k, 17 + 2 * q:18 + 2 * q])} "
                          f"PT {rel(t[3 + g, k, 18 + 2 * q:19 + 2 * q])} done {rel(t[3 + g, k, 4 + 4 * q:5 + 4 * q])}")
