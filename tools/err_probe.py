import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import golfer_b200
from oracle import segnet as osegnet
cfg = golfer_b200.V0
params = golfer_b200.params.make_params(cfg, 1234)
for (B, T, seed) in ((4, 300, 0), (8, 64, 5)):
    skel = osegnet.synth_skeletons(B, T, cfg, seed=seed)
    want = osegnet.segment_ref(cfg, params, skel)
    seg = golfer_b200.Segmenter(cfg, params, precision="bf16", device=0, max_B=B, max_T=T)
    got = seg.segment(torch.from_numpy(skel).cuda()).cpu().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    flips = (got.argmax(-1) != want.argmax(-1)).mean()
    print(f"B={B} T={T} FRONT_FFMA={os.environ.get('GOLFER_FRONT_FFMA')} rel err {err:.3e} label flips {flips:.4f}")
    seg.ctx.close()
