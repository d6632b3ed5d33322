"""One batch of the headline alignment workload (profiling target): python tools/prof_align.py [N] [Ta] [Tb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golfer_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Ta = int(sys.argv[2]) if len(sys.argv) > 2 else 300
Tb = int(sys.argv[3]) if len(sys.argv) > 3 else 300
g = torch.Generator().manual_seed(0)
a = (torch.randn(N, Ta, 17, 2, generator=g).cumsum(1) * 0.05).cuda()
b = (torch.randn(N, Tb, 17, 2, generator=g).cumsum(1) * 0.05).cuda()
for _ in range(2):
    cost, path, plen = golfer_b200.host.align_batch(a, b)
torch.cuda.synchronize()
print("ok", float(cost.sum()))
