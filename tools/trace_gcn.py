"""Dump the fused-GCN kernel's clock64 trace (CTA 0, first tiles) — GOLFER_TRACE_GCN=1.

roles: 0 producer (ev cb: box issue), 1 mma (ev q: MMA1 issue after d1_empty; 16+j: xa_full woke;
32+j: w_full woke), 2 convert (q: d1_full woke; 16+q: xa_empty woke; 32+q: chunk published),
3 gate (cb: x_full woke; 16+cb: x_ready), 4 epilogue (0: acc_full woke; 1: tile done)."""
import os
import sys

os.environ["GOLFER_TRACE_GCN"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import golfer_b200  # noqa: E402

B = int(os.environ.get("QB", "256"))
seg = golfer_b200.Segmenter(golfer_b200.V0, precision="bf16", max_B=B, max_T=300)
x = torch.randn(B, 300, 17, 3, device="cuda")
for _ in range(2):
    seg.segment(x)
torch.cuda.synchronize()
raw = seg.ctx.debug_read("gcn_trace", 8 * 5 * 6 * 64 * 8).view(np.uint64).reshape(8, 5, 6, 64)
names = ["prod", "mma", "conv", "gate", "epi"]
for blk in range(1, 6):
    t = raw[blk].astype(np.int64)
    t0 = t[t > 0].min() if (t > 0).any() else 0
    cin, c = golfer_b200.V0.block_io()[blk]
    nq = 3 * cin // 64
    print(f"=== block {blk} ({cin}->{c}), nq={nq}; cycles relative to first event")
    for tile in range(1, 4):
        rel = lambda a: [int(v - t0) if v > 0 else -1 for v in a]
        print(f" tile {tile}: prod box-issue {rel(t[0, tile, :cin // 64])}")
        print(f"          gate x_full {rel(t[3, tile, :cin // 64])} x_ready {rel(t[3, tile, 16:16 + cin // 64])}")
        print(f"          gate computed {rel(t[3, tile, 32:32 + cin // 64])} fenced {rel(t[3, tile, 36:36 + cin // 64])} store issued {rel(t[3, tile, 48:48 + cin // 64])} prev drained {rel(t[3, tile, 52:52 + cin // 64])}")
        print(f"          mma  MMA1   {rel(t[1, tile, :nq])}")
        print(f"          mma  xa_full{rel(t[1, tile, 16:16 + nq])}")
        print(f"          mma  w_full {rel(t[1, tile, 32:32 + nq])}")
        print(f"          conv d1_full{rel(t[2, tile, :nq])}")
        print(f"          conv xa_empt{rel(t[2, tile, 16:16 + nq])}")
        print(f"          conv publish{rel(t[2, tile, 32:32 + nq])}")
        print(f"          epi  woke, stored, drained {rel(t[4, tile, :3])}")
