"""Known-answer tests of the pose-adapter oracle (oracle/pose.py) and of the host-side parsers; CPU only."""
import json

import numpy as np

import golfer_b200
from oracle import pose as opose


def _kp(T=6):
    kp = np.zeros((T, 17, 3), np.float32)
    kp[..., 2] = 0.9
    kp[:, :, 0] = np.arange(17, dtype=np.float32)[None] * 4 + 100      # x
    kp[:, :, 1] = np.arange(17, dtype=np.float32)[None] * 2 + 50       # y
    kp[:, [5, 6], 1] = 10.0                                            # shoulders
    kp[:, [11, 12], 1] = 74.0                                          # hips: torso length = 64 at x offset
    kp[:, 5, 0], kp[:, 6, 0] = 120.0, 136.0                            # shoulder centre x = 128
    kp[:, 11, 0], kp[:, 12, 0] = 120.0, 136.0                          # hip centre x = 128
    return kp


def test_hip_centred_and_torso_scaled_exact_numbers():
    kp = _kp()
    out = opose.normalize_pose(kp)
    # hip centre (128, 74), torso length 64 (a power of two: every result is exact)
    assert np.array_equal(out[:, 11, :2], np.tile(np.float32([-0.125, 0.0]), (6, 1)))
    assert np.array_equal(out[:, 12, :2], np.tile(np.float32([0.125, 0.0]), (6, 1)))
    assert np.array_equal(out[:, 5, :2], np.tile(np.float32([-0.125, -1.0]), (6, 1)))
    assert np.array_equal(out[..., 2], kp[..., 2])


def test_translation_and_power_of_two_scale_invariance():
    kp = _kp()
    ref = opose.normalize_pose(kp)
    moved = kp.copy()
    moved[..., 0] += 256
    moved[..., 1] -= 32
    assert np.array_equal(opose.normalize_pose(moved), ref)          # integer coordinates: exact
    big = kp.copy()
    big[..., :2] *= 4
    assert np.array_equal(opose.normalize_pose(big), ref)


def test_low_score_joints_are_masked_and_do_not_move_others():
    kp = _kp()
    ref = opose.normalize_pose(kp)
    kp[2, 9, 2] = 0.1
    out = opose.normalize_pose(kp)
    assert np.array_equal(out[2, 9], np.zeros(3, np.float32))
    out[2, 9] = ref[2, 9]
    assert np.array_equal(out, ref)


def test_hip_centre_forward_fill_backfill_and_scale_fallback():
    kp = _kp(T=5)
    kp[:, :, 0] += np.arange(5, dtype=np.float32)[:, None] * 8         # the subject walks 8 px per frame
    kp[0, 11, 2] = 0.0                                                 # frame 0: hips invalid -> backfill from frame 1
    kp[3, 12, 2] = 0.0                                                 # frame 3: hips invalid -> carry frame 2
    out = opose.normalize_pose(kp)
    # nose (joint 0, x = 100 + 8t) relative to the centre actually used (128 + 8*{1,1,2,2,4}) / 64
    want = (np.float32([100, 108, 116, 124, 132]) - np.float32([136, 136, 144, 144, 160])) / np.float32(64)
    assert np.array_equal(out[:, 0, 0], want)
    assert np.array_equal(out[0, 11], np.zeros(3, np.float32)) and np.array_equal(out[3, 12], np.zeros(3, np.float32))
    # no frame with four valid torso joints: scale 1; no valid hips at all: centre (0,0)
    kp2 = _kp(T=3)
    kp2[:, 5, 2] = 0.0
    out2 = opose.normalize_pose(kp2)
    assert np.array_equal(out2[:, 0, 0], np.float32([100 - 128] * 3))
    kp2[:, 11, 2] = 0.0
    out3 = opose.normalize_pose(kp2)
    assert np.array_equal(out3[:, 0, :2], kp2[:, 0, :2])


def test_batch_equals_per_clip_and_empty_batch():
    kp = opose.synth_keypoints(3, 40, seed=5)
    out = opose.normalize_pose(kp)
    for b in range(3):
        assert np.array_equal(out[b], opose.normalize_pose(kp[b]))
    assert opose.normalize_pose(np.zeros((0, 4, 17, 3), np.float32)).shape == (0, 4, 17, 3)


def test_parsers_accept_the_pose_estimator_layouts(tmp_path):
    kp = opose.synth_keypoints(1, 4, seed=2)[0]
    flat = [kp[t].reshape(-1).tolist() for t in range(4)]
    assert np.array_equal(golfer_b200.pose.keypoints_from_frames(flat), kp)
    assert np.array_equal(golfer_b200.pose.keypoints_from_frames([{"keypoints": f} for f in flat]), kp)
    # several people per frame: the best-scoring detection wins; a frame without people becomes zeros
    frames = [{"people": [{"keypoints": flat[t], "score": 0.9}, {"keypoints": [0.0] * 51, "score": 0.2}]}
              for t in range(3)] + [{"people": []}]
    got = golfer_b200.pose.keypoints_from_frames(frames)
    assert np.array_equal(got[:3], kp[:3]) and not got[3].any()
    # COCO results list with a gap at frame 12 and a second track
    res = [{"image_id": 10 + t, "keypoints": flat[t], "score": 0.8, "track_id": 1} for t in (0, 1, 3)]
    res.append({"image_id": 11, "keypoints": [1.0] * 51, "score": 0.95, "track_id": 2})
    got = golfer_b200.pose.keypoints_from_coco_results(res, track_id=1)
    assert got.shape == (4, 17, 3) and np.array_equal(got[[0, 1, 3]], kp[[0, 1, 3]]) and not got[2].any()
    p = tmp_path / "kp.json"
    p.write_text(json.dumps({"results": res}))
    assert np.array_equal(golfer_b200.pose.load_keypoints_json(str(p), track_id=1), got)
    p.write_text(json.dumps(flat))
    assert np.array_equal(golfer_b200.pose.load_keypoints_json(str(p)), kp)


def test_golden_pose_and_phase_fixtures(golden_dir):
    import os
    from oracle import align as oalign, align_native
    g = np.load(os.path.join(golden_dir, "next_rows_small.npz"))
    assert np.array_equal(opose.normalize_pose(g["pose_kp"], 0.3), g["pose_out"])
    a, b, la, lb = g["phase_a"], g["phase_b"], g["phase_la"], g["phase_lb"]
    for tag, pen in (("soft", 0.5), ("hard", np.inf)):
        cost, path, plen = align_native.align_phase_batch_c(a, b, la, lb, pen)
        assert np.array_equal(cost, g[f"phase_{tag}_cost"])
        assert np.array_equal(path, g[f"phase_{tag}_path"]) and np.array_equal(plen, g[f"phase_{tag}_plen"])
        for n in range(len(a)):
            c, p = oalign.align_phase_ref(a[n], b[n], la[n], lb[n], pen)
            assert c == g[f"phase_{tag}_cost"][n] and np.array_equal(p, g[f"phase_{tag}_path"][n, :len(p)])
