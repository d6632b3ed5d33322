"""golfer-b200: B200-native skeleton-sequence hot path (segment + align).

Public surface (BASELINE.json north_star):
    segment(skel[B,T,V,C]) -> logits[B,T,K]
    align(a, b) -> (cost, path)
plus the rows SURVEY.md 8f marks next: compare(a, b, path), normalize_pose(keypoints),
align_phase(a, b, labels_a, labels_b, penalty), EmbedAligner(encoder_blob).align(a, b) (learned alignment embedding).
Everything computes in hand-written sm_100a CUDA kernels behind the C ABI in
include/golfer_b200.h; there is no CPU fallback.
"""
from . import config, params, pose, shard  # noqa: F401
from .config import V0, V0_STRESS, GolfSegConfig  # noqa: F401
from .host import (  # noqa: F401
    EmbedAligner,
    GolferError,
    Segmenter,
    align,
    align_phase,
    compare,
    library_path,
    load_library,
    normalize_pose,
    segment,
)
