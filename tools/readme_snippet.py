import sys, os
sys.path.insert(0, os.getcwd())
import torch, golfer_b200
kp = torch.rand(8, 300, 17, 3, device="cuda") * 400
skel = golfer_b200.normalize_pose(kp, min_score=0.3)
seg = golfer_b200.Segmenter(golfer_b200.V0, precision="bf16", max_B=8, max_T=300)
logits, labels = seg.segment(skel, return_labels=True)
a, b = skel[0::2, :, :, :2].contiguous(), skel[1::2, :, :, :2].contiguous()
cost, path = golfer_b200.align(a, b)
cost_p, path_p, plen = golfer_b200.align_phase(a, b, labels[0::2].contiguous(), labels[1::2].contiguous(), 0.5)
dist = golfer_b200.compare(a, b, path_p, plen)
print(logits.shape, labels.shape, cost.shape, path.shape, cost_p.shape, path_p.shape, plen.tolist(), dist.shape, bool((cost_p >= cost).all()))
