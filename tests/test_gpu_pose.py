"""GPU parity for the pose input adapter, through the C ABI (bit-exact bar)."""
import numpy as np
import pytest
import torch

import golfer_b200
from oracle import pose as opose

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T,drop", [(3, 300, 0.1), (2, 1, 0.0), (4, 37, 0.5), (1, 1800, 0.05), (2, 64, 1.0)])
def test_matches_oracle_bitwise(B, T, drop):
    kp = opose.synth_keypoints(B, T, seed=B * 1000 + T, drop=drop)
    want = opose.normalize_pose(kp, 0.3)
    got = golfer_b200.normalize_pose(torch.from_numpy(kp).cuda(), 0.3).cpu().numpy()
    assert np.array_equal(got, want)
    # host arrays in, host arrays out; single clip without the batch axis
    assert np.array_equal(golfer_b200.normalize_pose(kp[0], 0.3), want[0])


def test_edge_cases_bitwise():
    kp = opose.synth_keypoints(4, 50, seed=9, drop=0.0)
    kp[0, :10, 11, 2] = 0.0            # leading frames without valid hips: backfill
    kp[1, 20:30, 12, 2] = 0.0          # a gap: forward fill
    kp[2, :, 5, 2] = 0.0               # never a valid torso: scale 1
    kp[3, :, [11, 12], 2] = 0.0        # never valid hips: centre (0,0)
    for thr in (0.3, 0.0, 2.0):
        want = opose.normalize_pose(kp, thr)
        got = golfer_b200.normalize_pose(torch.from_numpy(kp).cuda(), thr).cpu().numpy()
        assert np.array_equal(got, want), thr


def test_feeds_the_segmenter_and_rejects_bad_shapes():
    kp = opose.synth_keypoints(2, 40, seed=3)
    skel = golfer_b200.normalize_pose(torch.from_numpy(kp).cuda())
    logits = golfer_b200.segment(skel)
    assert logits.shape == (2, 40, golfer_b200.V0.num_classes) and torch.isfinite(logits).all()
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.normalize_pose(torch.zeros(2, 4, 17, 2).cuda())
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.normalize_pose(torch.zeros(2, 4, 8, 3).cuda())
    assert golfer_b200.normalize_pose(torch.zeros(0, 4, 17, 3).cuda()).shape == (0, 4, 17, 3)


def test_golden_fixtures_pose_and_phase(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "next_rows_small.npz"))
    got = golfer_b200.normalize_pose(torch.from_numpy(g["pose_kp"]).cuda(), 0.3).cpu().numpy()
    assert np.array_equal(got, g["pose_out"])
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for tag, pen in (("soft", 0.5), ("hard", float("inf"))):
        cost, path, plen = golfer_b200.align_phase(dev(g["phase_a"]), dev(g["phase_b"]), dev(g["phase_la"]),
                                                   dev(g["phase_lb"]), pen)
        assert np.array_equal(cost.cpu().numpy(), g[f"phase_{tag}_cost"])
        assert np.array_equal(path.cpu().numpy(), g[f"phase_{tag}_path"])
        assert np.array_equal(plen.cpu().numpy(), g[f"phase_{tag}_plen"])
