"""Clock trace of the fused temporal kernel (CTA 0, steps 20..27).  The trace points are compiled in only with
-DGOLFER_TCN_TRACE (they cost ~10 % of the kernel):
    GOLFER_NVCC_EXTRA=-DGOLFER_TCN_TRACE python <package>/build.py --force && python tools/trace_tcn.py [block]
and rebuild without the define afterwards."""
import os, sys
os.environ["GOLFER_TRACE_TCN"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import golfer_b200
B, T = 256, 300
seg = golfer_b200.Segmenter(golfer_b200.V0, seed=1234, precision="bf16", max_B=B, max_T=T)
x = torch.randn(B, T, 17, 3).cuda()
seg.segment(x); seg.segment(x)
torch.cuda.synchronize()
tr = seg.ctx.debug_read("tcn_trace", 8 * 4 * 8 * 16 * 8).view(np.uint64).reshape(8, 4, 8, 16).astype(np.int64)
names = {0: "epi leader", 1: "epi warp17", 2: "issuer"}
for blk in ([int(sys.argv[1])] if len(sys.argv) > 1 else [1, 3, 5]):
    t = tr[blk]
    base = t[0, 0, 0]
    print(f"=== block {blk}: times relative to the leader's step-20 loop top")
    for role in (0, 1, 2):
        print(names[role])
        for s in range(8):
            row = [(int(v - base) if v else -1) for v in t[role, s]]
            print("  step", 20 + s, row[:13])
    print("  leader iteration periods:", np.diff(t[0, :, 0]).tolist())
