"""CPU ORACLE (test infrastructure, never shipped or timed as the product):
fp32 PyTorch restatement of the swing-phase segmentation network.

PARITY UNPINNED: the reference ships no code, tests or golden vectors
(SURVEY.md section 0 / 8c).  This module is the DEFINITION of correct for
`segment`, written from the README's headings plus public literature defaults;
every size is an [ASSUMPTION] frozen in `golfer_b200.config.GolfSegConfig`.

Reference evidence each piece follows (file:line = /root/reference/README.md):
  GraphConv            README.md:27-28   "Spatial Module - Graph Convolution"
  MultiBranchTCN       README.md:29-30   "Temporal Module - Multi-branch Temporal Convolution"
  ChannelAttention     README.md:31-32   "Channel Attention"
  STJointAttention     README.md:33-34   "ST-Joint Attention"
  SegNet (block order) README.md:27-34   heading order GCN -> TCN -> CA -> STJA
  head / logits        README.md:17-18   action-segmentation stage

Layout is channels-last `[B,T,V,C]` everywhere.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference legs may
import this file.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import golfer_b200

GolfSegConfig = golfer_b200.config.GolfSegConfig


def _t(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class EvalBN(nn.Module):
    """Eval-mode BatchNorm over the last axis with explicit running statistics."""

    def __init__(self, p: Dict[str, np.ndarray], prefix: str, eps: float):
        super().__init__()
        for k in ("gamma", "beta", "mean", "var"):
            self.register_buffer(k, _t(p[f"{prefix}.{k}"]))
        self.eps = eps

    def forward(self, x):
        return (x - self.mean) / torch.sqrt(self.var + self.eps) * self.gamma + self.beta


class GraphConv(nn.Module):
    """Y = ReLU(BN(sum_p (A_p . X) W_p + b))   (README.md:27-28)."""

    def __init__(self, p, b, eps):
        super().__init__()
        self.A = nn.Parameter(_t(p[f"{b}.gcn.A"]), requires_grad=False)    # [P,V,V]
        self.W = nn.Parameter(_t(p[f"{b}.gcn.W"]), requires_grad=False)    # [P,Cin,C]
        self.b = nn.Parameter(_t(p[f"{b}.gcn.b"]), requires_grad=False)
        self.bn = EvalBN(p, f"{b}.gcn.bn", eps)

    def forward(self, x):                                   # [B,T,V,Cin]
        xa = torch.einsum("pwv,btvc->btpwc", self.A, x)     # adjacency contraction
        y = torch.einsum("btpwc,pcd->btwd", xa, self.W) + self.b
        return F.relu(self.bn(y))


class MultiBranchTCN(nn.Module):
    """R branches of (1x1 -> BN -> ReLU -> 3-tap dilated conv over T -> BN), concat
    (README.md:29-30).  Zero padding in time, no stride."""

    def __init__(self, p, b, cfg: GolfSegConfig):
        super().__init__()
        self.W1 = nn.Parameter(_t(p[f"{b}.tcn.W1"]), requires_grad=False)   # [C,C]
        self.b1 = nn.Parameter(_t(p[f"{b}.tcn.b1"]), requires_grad=False)
        self.bn1 = EvalBN(p, f"{b}.tcn.bn1", cfg.bn_eps)
        self.W2 = nn.Parameter(_t(p[f"{b}.tcn.W2"]), requires_grad=False)   # [R,k,C/R,C/R]
        self.b2 = nn.Parameter(_t(p[f"{b}.tcn.b2"]), requires_grad=False)
        self.bn2 = EvalBN(p, f"{b}.tcn.bn2", cfg.bn_eps)
        self.dil = cfg.dilations

    def forward(self, y):                                   # [B,T,V,C]
        B, T, V, C = y.shape
        R = self.W2.shape[0]
        cr = C // R
        h = F.relu(self.bn1(y @ self.W1 + self.b1))
        outs = []
        for r in range(R):
            hr = h[..., r * cr:(r + 1) * cr]
            d = self.dil[r]
            hp = F.pad(hr, (0, 0, 0, 0, d, d))              # zero-pad the T axis
            z = 0
            for j in range(3):                              # tap j reads frame t + (j-1)*d
                z = z + hp[:, j * d:j * d + T] @ self.W2[r, j]
            outs.append(z)
        z = torch.cat(outs, dim=-1) + self.b2
        return self.bn2(z)


class ChannelAttention(nn.Module):
    """SE gate: mean over (T,V) -> FC -> ReLU -> FC -> sigmoid (README.md:31-32)."""

    def __init__(self, p, b):
        super().__init__()
        for k in ("W1", "b1", "W2", "b2"):
            setattr(self, k, nn.Parameter(_t(p[f"{b}.se.{k}"]), requires_grad=False))

    def gate(self, u):
        m = u.mean(dim=(1, 2))                              # [B,C]
        return torch.sigmoid(F.relu(m @ self.W1 + self.b1) @ self.W2 + self.b2)

    def forward(self, u):
        return u * self.gate(u)[:, None, None, :]


def hardswish(x):
    return x * torch.clamp(x + 3.0, 0.0, 6.0) / 6.0


class STJointAttention(nn.Module):
    """Factorised frame x joint gate (README.md:33-34; EfficientGCN-style ST-JointAtt
    recalled from the literature)."""

    def __init__(self, p, b, eps):
        super().__init__()
        for k in ("W", "b", "Wt", "bt", "Wv", "bv"):
            setattr(self, k, nn.Parameter(_t(p[f"{b}.stj.{k}"]), requires_grad=False))
        self.bn = EvalBN(p, f"{b}.stj.bn", eps)

    def gates(self, x):
        T = x.shape[1]
        xt = x.mean(dim=2)                                  # [B,T,C]
        xv = x.mean(dim=1)                                  # [B,V,C]
        cat = torch.cat([xt, xv], dim=1)                    # [B,T+V,C]
        att = hardswish(self.bn(cat @ self.W + self.b))     # [B,T+V,C/j]
        at = torch.sigmoid(att[:, :T] @ self.Wt + self.bt)  # [B,T,C]
        av = torch.sigmoid(att[:, T:] @ self.Wv + self.bv)  # [B,V,C]
        return at, av

    def forward(self, x):
        at, av = self.gates(x)
        return x * (at[:, :, None, :] * av[:, None, :, :])


class Block(nn.Module):
    def __init__(self, p, i, cin, c, cfg):
        super().__init__()
        b = f"b{i}"
        self.gcn = GraphConv(p, b, cfg.bn_eps)
        self.tcn = MultiBranchTCN(p, b, cfg)
        self.has_res = cin != c
        if self.has_res:
            self.Wr = nn.Parameter(_t(p[f"{b}.res.W"]), requires_grad=False)
            self.br = nn.Parameter(_t(p[f"{b}.res.b"]), requires_grad=False)
            self.bnr = EvalBN(p, f"{b}.res.bn", cfg.bn_eps)
        self.se = ChannelAttention(p, b)
        self.stj = STJointAttention(p, b, cfg.bn_eps)

    def pre_attention(self, x):
        res = self.bnr(x @ self.Wr + self.br) if self.has_res else x
        return F.relu(self.tcn(self.gcn(x)) + res)

    def forward(self, x):
        return self.stj(self.se(self.pre_attention(x)))


class SegNet(nn.Module):
    """skel[B,T,V,Cin] -> logits[B,T,K]   (SURVEY 8a row a6)."""

    def __init__(self, cfg: GolfSegConfig, p: Dict[str, np.ndarray]):
        super().__init__()
        self.cfg = cfg
        self.data_bn = EvalBN(p, "data_bn", cfg.bn_eps)
        self.blocks = nn.ModuleList(
            [Block(p, i, cin, c, cfg) for i, (cin, c) in enumerate(cfg.block_io())])
        self.Wh = nn.Parameter(_t(p["head.W"]), requires_grad=False)
        self.bh = nn.Parameter(_t(p["head.b"]), requires_grad=False)
        self.eval()

    def forward(self, skel, return_features: bool = False):
        B, T, V, C = skel.shape
        x = self.data_bn(skel.reshape(B, T, V * C)).reshape(B, T, V, C)
        feats = []
        for blk in self.blocks:
            x = blk(x)
            if return_features:
                feats.append(x)
        logits = x.mean(dim=2) @ self.Wh + self.bh
        return (logits, feats) if return_features else logits


@torch.no_grad()
def segment_ref(cfg: GolfSegConfig, params: Dict[str, np.ndarray], skel) -> np.ndarray:
    """fp32 oracle: logits [B,T,K] as numpy float32."""
    net = SegNet(cfg, params)
    x = torch.as_tensor(np.asarray(skel), dtype=torch.float32)
    return net(x).numpy()


def labels_from_logits(logits: np.ndarray) -> np.ndarray:
    """Per-frame phase label = first maximum over K (ties -> lowest class index)."""
    return np.argmax(logits, axis=-1).astype(np.uint8)


def synth_skeletons(B: int, T: int, cfg: GolfSegConfig, seed: int = 0) -> np.ndarray:
    """Synthetic clips of SURVEY 8d: x,y ~ N(0,1) hip-centred, conf ~ U(0,1)."""
    rng = np.random.default_rng(seed)
    V = cfg.num_joints
    xy = rng.standard_normal((B, T, V, 2)).astype(np.float32)
    hip = 0.5 * (xy[:, :, 11:12] + xy[:, :, 12:13])
    xy = xy - hip
    conf = rng.uniform(0.0, 1.0, (B, T, V, 1)).astype(np.float32)
    out = np.concatenate([xy, conf], axis=-1)
    if cfg.in_channels != 3:
        out = np.resize(out, (B, T, V, cfg.in_channels))
    return np.ascontiguousarray(out, dtype=np.float32)


@torch.no_grad()
def spread_head_params(cfg: GolfSegConfig, params: Dict[str, np.ndarray], calib_seed: int = 99, gain: float = 6.0):
    """`v0-spread`: a copy of `params` whose head is re-centred and re-scaled so that per-frame labels spread
    over the classes with O(1) margins (option (a)).  With the plain random-init head 95 %
    of the frames of a synthetic clip fall into one class, so label parity would exercise almost no decision
    boundary.  Only head.W / head.b change (the network body, hence every kernel's work, is v0's); the
    calibration runs the oracle on 4 seeded clips: per class k the logit is standardised over those frames
    (z_k = (l_k - mean_k) / std_k) and multiplied by `gain`, expressed as new head weights:
        W'[:, k] = W[:, k] * gain / std_k,   b'[k] = (b[k] - mean_k) * gain / std_k."""
    net = SegNet(cfg, params)
    x = torch.from_numpy(synth_skeletons(4, 96, cfg, seed=calib_seed))
    logits = net(x).reshape(-1, cfg.num_classes)
    mean, std = logits.mean(0).numpy(), logits.std(0).numpy()
    out = dict(params)
    scale = (gain / std).astype(np.float32)
    out["head.W"] = (params["head.W"] * scale[None, :]).astype(np.float32)
    out["head.b"] = ((params["head.b"] - mean) * scale).astype(np.float32)
    return out
