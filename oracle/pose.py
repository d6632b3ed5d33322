"""CPU ORACLE (test infrastructure, never shipped or timed as the product):
NumPy fp32 restatement of the input adapter between pose estimation and the
segmentation network.

PARITY UNPINNED: the reference ships no code; the only
evidence is README.md:15 ("Pose Estimation" produces the 2-D keypoints the
skeleton models consume).  This module is the DEFINITION of correct for
`normalize_pose`.  The normalisation itself (hip-centred, torso-scaled, low-score
joints masked) is the documented assumption of SURVEY.md 8d / 8f.4.

Arithmetic contract (what the CUDA kernel reproduces bit-for-bit); every operation
an individually rounded IEEE fp32 op, no fused multiply-add:
  per frame t (COCO-17 order: 5,6 shoulders, 11,12 hips)
      hip_t      = 0.5 * (kp[11] + kp[12])              (x and y)
      shoulder_t = 0.5 * (kp[5] + kp[6])
      len_t      = sqrt(dx*dx + dy*dy),  d = shoulder_t - hip_t
      hip_ok_t   = score[11] >= min_score and score[12] >= min_score
      len_ok_t   = hip_ok_t and score[5] >= min_score and score[6] >= min_score
  centre_t = hip_t if hip_ok_t, else the centre of the latest earlier frame with
             hip_ok; frames before the first such frame use the first one; (0,0) if none
  scale    = (sum of len_t over len_ok frames, t ascending, sequential) / count,
             1 if there is no such frame or the mean is not > 0
  out[t,v] = ((x - centre_t.x) / scale, (y - centre_t.y) / scale, score)
             if score >= min_score else (0, 0, 0)
"""
from __future__ import annotations

import numpy as np

L_SHOULDER, R_SHOULDER, L_HIP, R_HIP = 5, 6, 11, 12


def normalize_pose(kp: np.ndarray, min_score: float = 0.3) -> np.ndarray:
    """kp [T,V,3] or [B,T,V,3] fp32 (x, y, score) -> normalised skeletons, same shape."""
    kp = np.asarray(kp, dtype=np.float32)
    if kp.ndim == 4:
        return np.stack([normalize_pose(k, min_score) for k in kp]) if len(kp) else kp.copy()
    T, V, C = kp.shape
    assert C == 3 and V > R_HIP
    thr = np.float32(min_score)
    half = np.float32(0.5)
    x, y, s = kp[..., 0], kp[..., 1], kp[..., 2]
    hx = half * (x[:, L_HIP] + x[:, R_HIP])
    hy = half * (y[:, L_HIP] + y[:, R_HIP])
    sx = half * (x[:, L_SHOULDER] + x[:, R_SHOULDER])
    sy = half * (y[:, L_SHOULDER] + y[:, R_SHOULDER])
    dx, dy = sx - hx, sy - hy
    q = dx * dx
    q = q + dy * dy
    length = np.sqrt(q)
    hip_ok = (s[:, L_HIP] >= thr) & (s[:, R_HIP] >= thr)
    len_ok = hip_ok & (s[:, L_SHOULDER] >= thr) & (s[:, R_SHOULDER] >= thr)
    # centres: forward fill, leading frames take the first valid one
    cx = np.zeros(T, np.float32)
    cy = np.zeros(T, np.float32)
    have = False
    lx = ly = np.float32(0)
    for t in range(T):
        if hip_ok[t]:
            if not have:
                cx[:t], cy[:t] = hx[t], hy[t]
                have = True
            lx, ly = hx[t], hy[t]
        cx[t], cy[t] = lx, ly
    acc = np.float32(0)
    n = 0
    for t in range(T):
        if len_ok[t]:
            acc = np.float32(acc + length[t])
            n += 1
    scale = np.float32(1)
    if n > 0:
        mean = np.float32(acc / np.float32(n))
        if mean > 0:
            scale = mean
    out = np.zeros_like(kp)
    keep = s >= thr
    out[..., 0] = np.where(keep, (x - cx[:, None]) / scale, np.float32(0))
    out[..., 1] = np.where(keep, (y - cy[:, None]) / scale, np.float32(0))
    out[..., 2] = np.where(keep, s, np.float32(0))
    return out


def synth_keypoints(B: int, T: int, V: int = 17, seed: int = 0, drop: float = 0.1) -> np.ndarray:
    """Image-space keypoints of a drifting, breathing skeleton with `drop` of the scores under 0.3."""
    g = np.random.default_rng(seed)
    base = g.normal(0, 40, (B, 1, V, 2)).astype(np.float32)
    base[:, :, [L_SHOULDER, R_SHOULDER], 1] -= 120          # shoulders above hips
    drift = np.cumsum(g.normal(0, 2, (B, T, 1, 2)), axis=1).astype(np.float32)
    jitter = g.normal(0, 1.5, (B, T, V, 2)).astype(np.float32)
    xy = np.float32(320) + base + drift + jitter
    score = g.uniform(0.3, 1.0, (B, T, V, 1)).astype(np.float32)
    low = g.uniform(0, 1, (B, T, V, 1)) < drop
    score = np.where(low, g.uniform(0, 0.3, (B, T, V, 1)).astype(np.float32), score).astype(np.float32)
    return np.concatenate([xy.astype(np.float32), score], -1)
