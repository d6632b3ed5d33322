"""CPU ORACLE (test infrastructure, never shipped or timed as the product): the LEARNED alignment embedding
(SURVEY.md 8f item 3).  The reference shows that its temporal-alignment model is trained (README.md:44-47: the
alignment section has a "Loss" plot) but ships neither the encoder nor the loss.

PARITY UNPINNED and an ASSUMPTION SET OF ITS OWN (`AlignEmbedConfig v0`, frozen here before any kernel existed):
  * per-frame encoder f: the (x, y) of the V = 17 joints, flattened in joint order (34 inputs)
        h = relu(x @ W1 + b1)      W1 [34, 128]
        f = h @ W2 + b2            W2 [128, 128]            D = 128, no normalisation
    seeded random weights (the loss that would train them is unknowable from the reference);
  * frame-to-frame cost in the Gram form, which IS the definition (it is what makes the cost a GEMM):
        c[i, j] = sqrt(max(|fa_i|^2 + |fb_j|^2 - 2 fa_i . fb_j, 0))
  * the DP, tie-break and backtrack of oracle/align.py unchanged.

Parity policy (declared up front): the CUDA path rounds the embeddings to bf16 (tensor-core
operands) and accumulates in fp32 in an order the hardware chooses, so it cannot be bit-exact against NumPy.
  * cost matrix: within 1e-2 relative of `embed_cost` (fp32), and within 1e-3 of `embed_cost(emulate_bf16=True)`,
    which rounds exactly where the kernel rounds (what is left is fp32 summation order in the encoder and the GEMM,
    and the occasional embedding element that sits on a bf16 rounding boundary and rounds the other way);
  * DTW total and path: BIT-EXACT against this module's DP run on the cost matrix the GPU produced (pins the DP
    and backtrack kernels), and the total within 1e-2 of the fp32 oracle's; agreement of the path with the fp32
    oracle's path is measured and reported, not asserted (near-ties move under the rounding).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import align as oalign

import golfer_b200

# sizes, the seeded weight generator and the blob packing are shared with the product (as for the segmentation net)
IN_DIM, HIDDEN, DIM = golfer_b200.params.EMBED_IN, golfer_b200.params.EMBED_HIDDEN, golfer_b200.params.EMBED_DIM
make_embed_params = golfer_b200.params.make_embed_params
pack_embed_blob = golfer_b200.params.pack_embed_blob


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what cvt.rn.bf16.f32 does to finite values)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def embed(frames: np.ndarray, p: Dict[str, np.ndarray]) -> np.ndarray:
    """frames [T, V, >=2] -> embeddings [T, 128] fp32."""
    T = frames.shape[0]
    x = np.ascontiguousarray(frames[..., :2], dtype=np.float32).reshape(T, -1)
    assert x.shape[1] == IN_DIM
    h = np.maximum(x @ p["W1"] + p["b1"], np.float32(0))
    return (h @ p["W2"] + p["b2"]).astype(np.float32)


def embed_cost(a: np.ndarray, b: np.ndarray, p: Dict[str, np.ndarray], emulate_bf16: bool = False) -> np.ndarray:
    """a [Ta,V,>=2], b [Tb,V,>=2] -> cost [Ta,Tb] fp32 (Gram form, module docstring)."""
    fa, fb = embed(a, p), embed(b, p)
    if emulate_bf16:
        fa, fb = bf16_round(fa), bf16_round(fb)
    na = np.sum(fa * fa, axis=1, dtype=np.float32)
    nb = np.sum(fb * fb, axis=1, dtype=np.float32)
    d2 = (na[:, None] + nb[None, :]) - np.float32(2) * (fa @ fb.T)
    return np.sqrt(np.maximum(d2, np.float32(0))).astype(np.float32)


def dtw_on_cost(c: np.ndarray):
    """(total fp32, path [L,2] int32) of oracle/align.py's DP on a given cost matrix."""
    D, dirs = oalign.dtw_accumulate(c)
    return np.float32(D[-1, -1]), oalign.dtw_backtrack(dirs)


def align_embed_ref(a: np.ndarray, b: np.ndarray, p: Dict[str, np.ndarray], emulate_bf16: bool = False):
    """One pair: (total cost, path, cost matrix)."""
    c = embed_cost(a, b, p, emulate_bf16)
    total, path = dtw_on_cost(c)
    return total, path, c
