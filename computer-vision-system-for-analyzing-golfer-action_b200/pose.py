"""Host-side input adapter: pose-estimation output files -> keypoint arrays for `normalize_pose`
(README.md:15 names pose estimation as the producer of the 2-D keypoints).

Only parsing lives here (no arithmetic): the normalisation itself runs on the GPU behind
`gs_normalize_pose`.  Accepted inputs are the two layouts COCO-17 pose estimators write:

* a COCO results list: one dict per detection with "image_id" (frame number) and "keypoints"
  = 51 numbers (x, y, score) x 17, optionally "score" (detection score) and "track_id";
* a per-frame list: element t is either the 51-number list itself, a dict with "keypoints", or a
  dict with "people"/"annotations"/"instances" holding such dicts (the best-scoring one is taken).
"""
from __future__ import annotations

import json
from typing import Any, Iterable, List, Optional, Sequence

import numpy as np

NUM_JOINTS = 17


def _as_keypoints(obj: Any) -> Optional[np.ndarray]:
    """One detection -> [17,3] fp32, or None if `obj` holds no keypoints."""
    if isinstance(obj, dict):
        obj = obj.get("keypoints")
    if obj is None:
        return None
    a = np.asarray(obj, dtype=np.float32)
    if a.size != NUM_JOINTS * 3:
        raise ValueError(f"expected {NUM_JOINTS * 3} keypoint numbers (x, y, score per COCO joint), got {a.size}")
    return a.reshape(NUM_JOINTS, 3)


def _det_score(d: Any) -> float:
    if isinstance(d, dict) and "score" in d:
        return float(d["score"])
    k = _as_keypoints(d)
    return float(k[:, 2].mean()) if k is not None else -1.0


def _best(dets: Sequence[Any], track_id: Optional[int]) -> Optional[np.ndarray]:
    if track_id is not None:
        dets = [d for d in dets if isinstance(d, dict) and d.get("track_id") == track_id]
    dets = [d for d in dets if _as_keypoints(d) is not None]
    if not dets:
        return None
    return _as_keypoints(max(dets, key=_det_score))


def keypoints_from_frames(frames: Iterable[Any], track_id: Optional[int] = None) -> np.ndarray:
    """Per-frame records -> [T,17,3] fp32.  A frame without a usable detection becomes all zeros
    (score 0: `normalize_pose` masks it and carries the previous hip centre over it)."""
    rows: List[np.ndarray] = []
    for fr in frames:
        k = None
        if isinstance(fr, dict):
            for key in ("people", "annotations", "instances"):
                if key in fr:
                    k = _best(fr[key], track_id)
                    break
            else:
                k = _as_keypoints(fr)
        elif fr is not None and len(fr):
            first = fr[0]
            if isinstance(first, (dict, list, tuple, np.ndarray)):
                k = _best(fr, track_id)           # a list of detections
            else:
                k = _as_keypoints(fr)             # the 51 numbers themselves
        rows.append(k if k is not None else np.zeros((NUM_JOINTS, 3), np.float32))
    if not rows:
        return np.zeros((0, NUM_JOINTS, 3), np.float32)
    return np.stack(rows).astype(np.float32)


def keypoints_from_coco_results(results: Sequence[dict], track_id: Optional[int] = None,
                                num_frames: Optional[int] = None) -> np.ndarray:
    """COCO keypoint-results list (one dict per detection, "image_id" = frame index) -> [T,17,3]."""
    by_frame: dict = {}
    for d in results:
        by_frame.setdefault(int(d["image_id"]), []).append(d)
    if not by_frame and not num_frames:
        return np.zeros((0, NUM_JOINTS, 3), np.float32)
    first = min(by_frame) if by_frame else 0
    T = num_frames if num_frames is not None else max(by_frame) - first + 1
    return keypoints_from_frames((by_frame.get(first + t, []) for t in range(T)), track_id)


def load_keypoints_json(path: str, track_id: Optional[int] = None) -> np.ndarray:
    """Read a JSON file in either accepted layout -> [T,17,3] fp32."""
    with open(path) as f:
        data = json.load(f)
    if isinstance(data, dict):
        for key in ("frames", "results", "annotations"):
            if key in data:
                data = data[key]
                break
    if data and isinstance(data[0], dict) and "image_id" in data[0]:
        return keypoints_from_coco_results(data, track_id)
    return keypoints_from_frames(data, track_id)
