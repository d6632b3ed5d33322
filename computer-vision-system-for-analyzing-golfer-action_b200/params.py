"""Parameters of the segmentation network: seeded generator, BN folding, blob packing.

No weights ship with the reference ("Checkpoint / resume"),
so the bench and the parity tests use seeded random-init parameters.  The raw
dictionary produced here is what BOTH the CPU oracle (`oracle/segnet.py`) and
the CUDA library consume, so they share parameters bit-for-bit.

All matrices are stored `[in, out]` (row-major, `out` contiguous): y = x @ W + b.

Raw names (numpy float32):
  data_bn.{gamma,beta,mean,var}                [V*Cin]   index v*Cin + c
  b{i}.gcn.A [P,V,V]  b{i}.gcn.W [P,Cin,C]  b{i}.gcn.b [C]  b{i}.gcn.bn.* [C]
  b{i}.tcn.W1 [C,C]   b{i}.tcn.b1 [C]  b{i}.tcn.bn1.* [C]      (branch r = out cols r*C/R..)
  b{i}.tcn.W2 [R,k,C/R,C/R]  b{i}.tcn.b2 [C]  b{i}.tcn.bn2.* [C]
  b{i}.res.W [Cin,C]  b{i}.res.b [C]  b{i}.res.bn.* [C]         (only when Cin != C)
  b{i}.se.W1 [C,C/s] b{i}.se.b1  b{i}.se.W2 [C/s,C] b{i}.se.b2
  b{i}.stj.W [C,C/j] b{i}.stj.b b{i}.stj.bn.* [C/j]
  b{i}.stj.Wt [C/j,C] b{i}.stj.bt  b{i}.stj.Wv [C/j,C] b{i}.stj.bv
  head.W [C,K]  head.b [K]
"""
from __future__ import annotations

import hashlib
from typing import Dict

import numpy as np

from .config import COCO_CENTER, COCO_EDGES, GolfSegConfig

BLOB_MAGIC = 0x30575347  # "GSW0" little-endian


def build_adjacency(cfg: GolfSegConfig) -> np.ndarray:
    """[P,V,V] spatial-partition adjacency (self / centripetal / centrifugal),
    column-normalised, ST-GCN "spatial" strategy written from the published
    description (README.md:27 names the module only).

    A[p, w, v] multiplies input joint v into output joint w.
    """
    V = cfg.num_joints
    if cfg.num_partitions != 3:
        raise ValueError("v0 defines the 3-partition spatial strategy only")
    adj = np.zeros((V, V), dtype=np.float64)
    for i, j in COCO_EDGES:
        adj[i, j] = adj[j, i] = 1.0
    # hop distance from the centre joint (BFS over the tree)
    hop = np.full(V, -1, dtype=np.int64)
    hop[COCO_CENTER] = 0
    frontier = [COCO_CENTER]
    while frontier:
        nxt = []
        for u in frontier:
            for w in range(V):
                if adj[u, w] and hop[w] < 0:
                    hop[w] = hop[u] + 1
                    nxt.append(w)
        frontier = nxt
    full = adj + np.eye(V)
    norm = full / full.sum(axis=0, keepdims=True)      # column-normalised
    A = np.zeros((3, V, V), dtype=np.float64)
    for w in range(V):
        for v in range(V):
            if not full[w, v]:
                continue
            if hop[v] == hop[w]:
                A[0, w, v] = norm[w, v]                # self / same distance
            elif hop[v] > hop[w]:
                A[1, w, v] = norm[w, v]                # v is further out: centripetal flow into w
            else:
                A[2, w, v] = norm[w, v]                # centrifugal
    return A.astype(np.float32)


def _bn(rng: np.random.Generator, n: int) -> Dict[str, np.ndarray]:
    return {
        "gamma": rng.uniform(0.8, 1.2, n).astype(np.float32),
        "beta": rng.normal(0.0, 0.1, n).astype(np.float32),
        "mean": rng.normal(0.0, 0.1, n).astype(np.float32),
        "var": rng.uniform(0.5, 1.5, n).astype(np.float32),
    }


def make_params(cfg: GolfSegConfig, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Seeded random-init parameters with randomised BN running statistics
    (so that folding is actually exercised)."""
    rng = np.random.default_rng(seed)
    V, P, R, k = cfg.num_joints, cfg.num_partitions, cfg.num_branches, cfg.kernel_size
    p: Dict[str, np.ndarray] = {}

    def put_bn(prefix, n):
        for key, val in _bn(rng, n).items():
            p[f"{prefix}.{key}"] = val

    def lin(fan_in, shape, gain):
        return (rng.standard_normal(shape) * (gain / np.sqrt(fan_in))).astype(np.float32)

    put_bn("data_bn", V * cfg.in_channels)
    A0 = build_adjacency(cfg)
    for i, (cin, c) in enumerate(cfg.block_io()):
        b = f"b{i}"
        # learnable adjacency: base graph x edge importance + small dense term, so
        # kernels must treat A as dense data.
        imp = rng.uniform(0.8, 1.2, A0.shape)
        dense = rng.normal(0.0, 0.01, A0.shape)
        p[f"{b}.gcn.A"] = (A0 * imp + dense).astype(np.float32)
        p[f"{b}.gcn.W"] = lin(P * cin, (P, cin, c), 2.0)
        p[f"{b}.gcn.b"] = rng.normal(0, 0.05, c).astype(np.float32)
        put_bn(f"{b}.gcn.bn", c)
        p[f"{b}.tcn.W1"] = lin(c, (c, c), np.sqrt(2.0))
        p[f"{b}.tcn.b1"] = rng.normal(0, 0.05, c).astype(np.float32)
        put_bn(f"{b}.tcn.bn1", c)
        cr = c // R
        p[f"{b}.tcn.W2"] = lin(k * cr, (R, k, cr, cr), np.sqrt(2.0))
        p[f"{b}.tcn.b2"] = rng.normal(0, 0.05, c).astype(np.float32)
        put_bn(f"{b}.tcn.bn2", c)
        if cin != c:
            p[f"{b}.res.W"] = lin(cin, (cin, c), 1.0)
            p[f"{b}.res.b"] = rng.normal(0, 0.05, c).astype(np.float32)
            put_bn(f"{b}.res.bn", c)
        cs = c // cfg.se_reduction
        p[f"{b}.se.W1"] = lin(c, (c, cs), np.sqrt(2.0))
        p[f"{b}.se.b1"] = rng.normal(0, 0.05, cs).astype(np.float32)
        p[f"{b}.se.W2"] = lin(cs, (cs, c), 1.0)
        p[f"{b}.se.b2"] = rng.normal(2.0, 0.5, c).astype(np.float32)
        cj = c // cfg.stj_reduction
        p[f"{b}.stj.W"] = lin(c, (c, cj), np.sqrt(2.0))
        p[f"{b}.stj.b"] = rng.normal(0, 0.05, cj).astype(np.float32)
        put_bn(f"{b}.stj.bn", cj)
        p[f"{b}.stj.Wt"] = lin(cj, (cj, c), 1.0)
        p[f"{b}.stj.bt"] = rng.normal(2.0, 0.5, c).astype(np.float32)
        p[f"{b}.stj.Wv"] = lin(cj, (cj, c), 1.0)
        p[f"{b}.stj.bv"] = rng.normal(2.0, 0.5, c).astype(np.float32)
    c = cfg.widths[-1]
    p["head.W"] = lin(c, (c, cfg.num_classes), 4.0)
    p["head.b"] = rng.normal(0, 0.1, cfg.num_classes).astype(np.float32)
    return p


def _fold(W: np.ndarray, b: np.ndarray, p: Dict[str, np.ndarray], bn: str, eps: float):
    """Fold eval-mode BatchNorm `bn` (over the LAST axis of W) into (W, b), in fp32."""
    scale = (p[f"{bn}.gamma"] / np.sqrt(p[f"{bn}.var"] + np.float32(eps))).astype(np.float32)
    Wf = (W * scale).astype(np.float32)
    bf = ((b - p[f"{bn}.mean"]) * scale + p[f"{bn}.beta"]).astype(np.float32)
    return Wf, bf


def fold_params(cfg: GolfSegConfig, p: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """BN-folded parameters in the order the blob stores them."""
    eps = cfg.bn_eps
    R = cfg.num_branches
    f: Dict[str, np.ndarray] = {}
    sc = (p["data_bn.gamma"] / np.sqrt(p["data_bn.var"] + np.float32(eps))).astype(np.float32)
    f["in.scale"] = sc
    f["in.shift"] = (p["data_bn.beta"] - p["data_bn.mean"] * sc).astype(np.float32)
    for i, (cin, c) in enumerate(cfg.block_io()):
        b = f"b{i}"
        f[f"{b}.A"] = p[f"{b}.gcn.A"]
        Wg, bg = _fold(p[f"{b}.gcn.W"], p[f"{b}.gcn.b"], p, f"{b}.gcn.bn", eps)
        f[f"{b}.Wg"] = Wg.reshape(cfg.num_partitions * cin, c)
        f[f"{b}.bg"] = bg
        f[f"{b}.W1"], f[f"{b}.b1"] = _fold(p[f"{b}.tcn.W1"], p[f"{b}.tcn.b1"], p, f"{b}.tcn.bn1", eps)
        cr = c // R
        s2 = (p[f"{b}.tcn.bn2.gamma"] / np.sqrt(p[f"{b}.tcn.bn2.var"] + np.float32(eps))).astype(np.float32)
        W2 = p[f"{b}.tcn.W2"] * s2.reshape(R, 1, 1, cr)
        f[f"{b}.W2"] = W2.astype(np.float32)
        f[f"{b}.b2"] = ((p[f"{b}.tcn.b2"] - p[f"{b}.tcn.bn2.mean"]) * s2 + p[f"{b}.tcn.bn2.beta"]).astype(np.float32)
        if cin != c:
            f[f"{b}.Wr"], f[f"{b}.br"] = _fold(p[f"{b}.res.W"], p[f"{b}.res.b"], p, f"{b}.res.bn", eps)
        for key in ("W1", "b1", "W2", "b2"):
            f[f"{b}.se.{key}"] = p[f"{b}.se.{key}"]
        f[f"{b}.stj.W"], f[f"{b}.stj.b"] = _fold(p[f"{b}.stj.W"], p[f"{b}.stj.b"], p, f"{b}.stj.bn", eps)
        for key in ("Wt", "bt", "Wv", "bv"):
            f[f"{b}.stj.{key}"] = p[f"{b}.stj.{key}"]
    f["head.W"] = p["head.W"]
    f["head.b"] = p["head.b"]
    return f


def pack_blob(cfg: GolfSegConfig, p: Dict[str, np.ndarray]) -> np.ndarray:
    """Host weight blob handed to `gs_create` (include/golfer_b200.h): a 4-word
    header {magic, n_floats, n_blocks, reserved} followed by the folded
    parameters as float32 in `fold_params` order."""
    f = fold_params(cfg, p)
    body = np.concatenate([np.ascontiguousarray(v, dtype=np.float32).ravel() for v in f.values()])
    head = np.array([BLOB_MAGIC, body.size, cfg.num_blocks, 0], dtype=np.uint32).view(np.float32)
    return np.concatenate([head, body])


def blob_sha256(blob: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(blob).tobytes()).hexdigest()


# ---- learned alignment embedding: AlignEmbedConfig v0 -------------------------------------
# [ASSUMPTION] per-frame MLP over the (x, y) of the 17 joints: 34 -> 128 (ReLU) -> 128.  The reference trains its
# alignment model (README.md:44-47 shows a loss curve) but ships neither encoder nor loss, so these are seeded
# random weights; the CPU oracle (oracle/embed.py) and the CUDA library consume the same dictionary.
EMBED_IN, EMBED_HIDDEN, EMBED_DIM = 34, 128, 128


def make_embed_params(seed: int = 4321) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    return {
        "W1": (rng.standard_normal((EMBED_IN, EMBED_HIDDEN)) * np.sqrt(2.0 / EMBED_IN)).astype(np.float32),
        "b1": rng.normal(0, 0.05, EMBED_HIDDEN).astype(np.float32),
        "W2": (rng.standard_normal((EMBED_HIDDEN, EMBED_DIM)) * np.sqrt(1.0 / EMBED_HIDDEN)).astype(np.float32),
        "b2": rng.normal(0, 0.05, EMBED_DIM).astype(np.float32),
    }


def pack_embed_blob(p: Dict[str, np.ndarray]) -> np.ndarray:
    """Host blob handed to gs_set_align_encoder: W1 [34,128], b1 [128], W2 [128,128], b2 [128], fp32, in this order."""
    return np.concatenate([np.ascontiguousarray(p[k], dtype=np.float32).ravel() for k in ("W1", "b1", "W2", "b2")])
