"""Two bf16 forwards of the headline batch (profiling target: the second one is warm).
python tools/prof_seg.py [B] [T] [stress]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golfer_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
cfg = golfer_b200.V0_STRESS if len(sys.argv) > 3 else golfer_b200.V0
seg = golfer_b200.Segmenter(cfg, seed=1234, precision="bf16", max_B=B, max_T=T)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, 17, 3, generator=g).cuda()
for _ in range(2):
    out = seg.segment(x)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
