"""CPU ORACLE (test infrastructure, never shipped or timed as the product): an INDEPENDENT
restatement of the 3-partition spatial adjacency (README.md:27 names the
module only), so that the product's `params.build_adjacency` is checked against something it
did not write itself.

PARITY UNPINNED: the reference ships no code.  [ASSUMPTION] ST-GCN "spatial"
partitioning strategy over the COCO-17 tree, centre joint 0, as frozen in GolfSegConfig v0.

Formulation (different from the product's BFS + per-pair loop on purpose):
  * hop distances from matrix powers of the adjacency: hop[i,j] = min k with (I + E)^k [i,j] > 0;
  * the normalised graph Ā = (E + I) D^-1 with D = column degree;
  * A_0 = Ā masked to pairs at the same distance from the centre, A_1 = Ā masked to pairs whose
    SOURCE joint v is farther from the centre than the target w, A_2 = the rest.
A[p, w, v] multiplies input joint v into output joint w.
"""
from __future__ import annotations

import numpy as np

# COCO-17 skeleton, written out independently of golfer_b200.config:
# ankles-knees-hips, hips-shoulders, wrists-elbows-shoulders, shoulders-nose, eyes-nose, ears-eyes.
EDGES_BY_LIMB = {
    "left leg": [(15, 13), (13, 11)],
    "right leg": [(16, 14), (14, 12)],
    "torso sides": [(11, 5), (12, 6)],
    "left arm": [(9, 7), (7, 5)],
    "right arm": [(10, 8), (8, 6)],
    "neck": [(5, 0), (6, 0)],
    "face": [(1, 0), (3, 1), (2, 0), (4, 2)],
}
CENTER = 0
V = 17


def hop_distance_matrix(num_joints: int = V) -> np.ndarray:
    """All-pairs hop distance by repeated squaring-free powers of (I + E)."""
    E = np.zeros((num_joints, num_joints), dtype=np.int64)
    for limb in EDGES_BY_LIMB.values():
        for i, j in limb:
            E[i, j] = E[j, i] = 1
    reach = np.eye(num_joints, dtype=np.int64)
    hop = np.full((num_joints, num_joints), -1, dtype=np.int64)
    hop[np.eye(num_joints, dtype=bool)] = 0
    step = np.eye(num_joints, dtype=np.int64) + E
    for k in range(1, num_joints):
        reach = (reach @ step > 0).astype(np.int64)
        newly = (reach > 0) & (hop < 0)
        hop[newly] = k
    assert (hop >= 0).all(), "skeleton graph is not connected"
    return hop


def spatial_partitions(num_joints: int = V) -> np.ndarray:
    """[3, V, V] float32: self / centripetal / centrifugal, column-normalised over the whole graph."""
    hop = hop_distance_matrix(num_joints)
    linked = (hop <= 1).astype(np.float64)              # E + I
    normalised = linked / linked.sum(axis=0, keepdims=True)
    d = hop[CENTER]                                     # distance of every joint from the centre
    dw, dv = d[:, None], d[None, :]                     # target w (rows), source v (columns)
    same = (dw == dv)
    source_farther = (dv > dw)
    A = np.stack([normalised * same, normalised * source_farther, normalised * ~(same | source_farther)])
    return A.astype(np.float32)
