"""Frozen architecture description of the swing-phase segmentation network.

The reference ships no code; the only evidence for the
network is the README's headings:

  * README.md:27  "Spatial Module - Graph Convolution"
  * README.md:29  "Temporal Module - Multi-branch Temporal Convolution"
  * README.md:31  "Channel Attention"
  * README.md:33  "ST-Joint Attention"
  * README.md:17  action-segmentation stage (per-frame phase labels)

Every size below is a documented ASSUMPTION (SURVEY.md section 8a,
"GolfSegConfig v0"), frozen before the first kernel was written.  Results
always carry `config_hash()`.
"""
from __future__ import annotations

import hashlib
import json
from dataclasses import asdict, dataclass
from typing import Tuple

# COCO-17 keypoint order: 0 nose, 1-2 eyes, 3-4 ears, 5-6 shoulders, 7-8 elbows,
# 9-10 wrists, 11-12 hips, 13-14 knees, 15-16 ankles.
COCO_EDGES: Tuple[Tuple[int, int], ...] = (
    (15, 13), (13, 11), (16, 14), (14, 12), (11, 5), (12, 6), (9, 7), (7, 5),
    (10, 8), (8, 6), (5, 0), (6, 0), (1, 0), (3, 1), (2, 0), (4, 2),
)
COCO_CENTER = 0


@dataclass(frozen=True)
class GolfSegConfig:
    """Sizes of the segmentation network (all [ASSUMPTION], see module docstring)."""

    version: str = "v0"
    num_joints: int = 17          # V, COCO-17
    in_channels: int = 3          # (x, y, confidence)
    num_partitions: int = 3       # P: self / centripetal / centrifugal
    widths: Tuple[int, ...] = (64, 64, 128, 128, 256, 256)
    num_branches: int = 4         # R temporal branches
    kernel_size: int = 3          # taps per temporal branch
    dilations: Tuple[int, ...] = (1, 2, 3, 4)
    se_reduction: int = 4         # channel attention bottleneck
    stj_reduction: int = 4        # ST-joint attention bottleneck
    num_classes: int = 9          # K: 8 swing phases + background
    bn_eps: float = 1e-5

    def __post_init__(self):
        if len(self.dilations) != self.num_branches:
            raise ValueError("one dilation per temporal branch")
        if self.kernel_size != 3:
            raise ValueError("only 3-tap temporal branches are defined in v0")
        for c in self.widths:
            if c % self.num_branches or c % self.se_reduction or c % self.stj_reduction:
                raise ValueError(f"width {c} not divisible by branches/reductions")

    @property
    def num_blocks(self) -> int:
        return len(self.widths)

    def block_io(self):
        """[(Cin, Cout)] per block."""
        cin = self.in_channels
        out = []
        for c in self.widths:
            out.append((cin, c))
            cin = c
        return out

    def config_hash(self) -> str:
        blob = json.dumps(asdict(self), sort_keys=True).encode()
        return hashlib.sha256(blob).hexdigest()[:12]

    # ---- algorithmic work, un-padded -----------------
    def flops_per_clip(self, T: int) -> float:
        V, P, R, k = self.num_joints, self.num_partitions, self.num_branches, self.kernel_size
        rows = T * V
        total = 0.0
        for cin, c in self.block_io():
            total += 2.0 * T * V * V * cin * P            # adjacency contraction
            total += 2.0 * rows * (P * cin) * c           # channel mix
            total += 2.0 * rows * c * c                   # branch 1x1 reduce
            total += R * 2.0 * rows * k * (c // R) ** 2   # dilated taps
            if cin != c:
                total += 2.0 * rows * cin * c             # residual projection
            cs, cj = c // self.se_reduction, c // self.stj_reduction
            total += 2.0 * 2 * c * cs                     # SE FCs
            total += 2.0 * (T + V) * c * cj * 2           # ST-joint FCs
        total += 2.0 * T * self.widths[-1] * self.num_classes
        return total

    def compulsory_bytes_per_clip(self, T: int, act_bytes: int = 2) -> float:
        """SURVEY.md 8d / BASELINE.md section 4 figure: every block reads its input
        and writes its output once (15.70 MB at T=300, bf16)."""
        rows = T * self.num_joints
        return float(sum((cin + c) * act_bytes * rows for cin, c in self.block_io()))


V0 = GolfSegConfig()
# BASELINE.json configs[4]: long-sequence stress, 8 temporal branches.
V0_STRESS = GolfSegConfig(version="v0-stress", num_branches=8, dilations=(1, 2, 3, 4, 5, 6, 7, 8))
