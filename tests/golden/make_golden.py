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
from oracle import align, segnet  # noqa: E402

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


if __name__ == "__main__":
    golden_align()
    golden_segnet()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
