"""Regenerate the committed golden fixtures from the CPU oracle.

    python tests/golden/make_golden.py

The reference holds no golden vectors (parity unpinned); these
fixtures pin the ORACLE (its arithmetic and the seeded parameter generator)
against drift across numpy / torch versions, and give the GPU parity tests
fixed small cases that do not need the oracle's Python to regenerate inputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import golfer_b200  # noqa: E402
from oracle import align, pose, segnet  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TINY = golfer_b200.GolfSegConfig(version="tiny", widths=(16, 16, 32))


def golden_align():
    out = {}
    for tag, (N, Ta, Tb, seed) in {"sq": (3, 16, 16, 11), "rect": (4, 12, 9, 12), "one": (2, 1, 7, 13)}.items():
        a, b = align.synth_swings(N, Ta, Tb, seed=seed)
        maxL = Ta + Tb - 1
        cm = np.stack([align.pair_cost(a[n], b[n]) for n in range(N)])
        cost = np.zeros(N, np.float32)
        path = np.full((N, maxL, 2), -1, np.int32)
        plen = np.zeros(N, np.int32)
        for n in range(N):
            c, p = align.align_ref(a[n], b[n])
            cost[n], plen[n] = c, len(p)
            path[n, :len(p)] = p
        out.update({f"{tag}_a": a, f"{tag}_b": b, f"{tag}_cm": cm, f"{tag}_cost": cost,
                    f"{tag}_path": path, f"{tag}_plen": plen})
    np.savez_compressed(os.path.join(HERE, "align_small.npz"), **out)


def golden_segnet():
    out = {}
    for tag, cfg, B, T in (("tiny", TINY, 2, 12), ("v0", golfer_b200.V0, 1, 24)):
        params = golfer_b200.params.make_params(cfg, 1234)
        blob = golfer_b200.params.pack_blob(cfg, params)
        skel = segnet.synth_skeletons(B, T, cfg, seed=3)
        logits = segnet.segment_ref(cfg, params, skel)
        out[f"{tag}_skel"] = skel
        out[f"{tag}_logits"] = logits
        out[f"{tag}_labels"] = segnet.labels_from_logits(logits)
        out[f"{tag}_blob_sha256"] = np.array(golfer_b200.params.blob_sha256(blob))
        out[f"{tag}_blob_sum"] = np.array(np.float64(blob[4:].astype(np.float64).sum()))
    np.savez_compressed(os.path.join(HERE, "segnet_small.npz"), **out)


def golden_next_rows():
    """pose adapter and phase-conditioned alignment."""
    out = {}
    kp = pose.synth_keypoints(3, 20, seed=21, drop=0.15)
    kp[1, :4, 11, 2] = 0.0          # leading frames without valid hips
    kp[2, :, 5, 2] = 0.0            # never a valid torso: scale 1
    out["pose_kp"] = kp
    out["pose_out"] = pose.normalize_pose(kp, 0.3)
    N, Ta, Tb = 3, 14, 11
    a, b = align.synth_swings(N, Ta, Tb, seed=31)
    rng = np.random.default_rng(5)
    la = np.sort(rng.integers(0, 4, (N, Ta)), axis=1).astype(np.uint8)
    lb = np.sort(rng.integers(0, 4, (N, Tb)), axis=1).astype(np.uint8)
    maxL = Ta + Tb - 1
    for tag, pen in (("soft", 0.5), ("hard", np.inf)):
        cost = np.zeros(N, np.float32)
        path = np.full((N, maxL, 2), -1, np.int32)
        plen = np.zeros(N, np.int32)
        for n in range(N):
            c, p = align.align_phase_ref(a[n], b[n], la[n], lb[n], pen)
            cost[n], plen[n] = c, len(p)
            path[n, :len(p)] = p
        out.update({f"phase_{tag}_cost": cost, f"phase_{tag}_path": path, f"phase_{tag}_plen": plen})
    out.update({"phase_a": a, "phase_b": b, "phase_la": la, "phase_lb": lb})
    np.savez_compressed(os.path.join(HERE, "next_rows_small.npz"), **out)


if __name__ == "__main__":
    golden_align()
    golden_segnet()
    golden_next_rows()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
