"""One small bf16 forward (debug helper): python tools/quick_seg.py [B] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import golfer_b200
from oracle import segnet as osegnet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cfg = golfer_b200.V0
if os.environ.get("WIDTHS"):
    cfg = golfer_b200.GolfSegConfig(version="dbg", widths=tuple(int(w) for w in os.environ["WIDTHS"].split(",")))
params = golfer_b200.params.make_params(cfg, 1234)
skel = osegnet.synth_skeletons(B, T, cfg, seed=3)
seg = golfer_b200.Segmenter(cfg, params, precision="bf16", max_B=B, max_T=T)
x = torch.from_numpy(skel).cuda()
upto = int(os.environ.get("UPTO", "-1"))
if upto >= 0:
    net = osegnet.SegNet(cfg, params)
    with torch.no_grad():
        _, feats = net(torch.from_numpy(skel), return_features=True)
    for i in range(upto + 1):
        got = seg.features(x, i).cpu().numpy()
        want = feats[i].numpy()
        print(f"block {i}: rel err {np.abs(got - want).max() / np.abs(want).max():.3e}")
else:
    got = seg.segment(x).cpu().numpy()
    want = osegnet.segment_ref(cfg, params, skel)
    print("rel err", np.abs(got - want).max() / np.abs(want).max())
