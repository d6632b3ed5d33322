"""Oracle self-consistency for the segmentation path (CPU only; SURVEY.md section 4 tier 1)."""
import os

import numpy as np
import torch

import golfer_b200
from oracle import segnet

TINY = golfer_b200.GolfSegConfig(version="tiny", widths=(16, 16, 32))


def _params(cfg=TINY, seed=1234):
    return golfer_b200.params.make_params(cfg, seed)


def test_config_work_figures_match_survey():
    cfg = golfer_b200.V0
    assert abs(cfg.flops_per_clip(300) / 1e9 - 7.772) < 1e-3
    assert abs(cfg.compulsory_bytes_per_clip(300) / 1e6 - 15.70) < 1e-2
    assert abs(golfer_b200.V0_STRESS.flops_per_clip(1800) / 1e9 - 42.67) < 1e-2
    assert abs(golfer_b200.V0_STRESS.compulsory_bytes_per_clip(1800) / 1e6 - 94.2) < 0.05


def test_adjacency_partitions():
    A = golfer_b200.params.build_adjacency(golfer_b200.V0)
    assert A.shape == (3, 17, 17)
    full = A.sum(0)
    assert np.allclose(full.sum(axis=0), 1.0, atol=1e-6)               # column-normalised
    assert np.count_nonzero(full) == 17 + 2 * 16                       # tree + self loops
    assert np.count_nonzero(A[0]) == 17                                # self partition = diagonal
    assert np.allclose(np.diag(A[0]), np.diag(full))
    assert np.count_nonzero(A[1]) == 16 and np.count_nonzero(A[2]) == 16
    assert not np.any((A[1] != 0) & (A[2] != 0))


def test_gcn_with_identity_adjacency_is_pointwise_conv():
    cfg = TINY
    p = _params()
    V = cfg.num_joints
    A = np.zeros((3, V, V), np.float32)
    A[0] = np.eye(V)
    p["b0.gcn.A"] = A
    g = segnet.GraphConv(p, "b0", cfg.bn_eps)
    x = torch.randn(2, 5, V, cfg.in_channels)
    y = g(x)
    ref = torch.relu(g.bn(x @ g.W[0] + g.b))
    assert torch.allclose(y, ref, atol=1e-6)


def test_tcn_zero_padding_and_dilation():
    cfg = TINY
    p = _params()
    t = segnet.MultiBranchTCN(p, "b1", cfg)
    C = 16
    y = torch.randn(1, 9, 17, C)
    out = t(y)
    # frame 0 of branch r must not see frames < 0: recompute by hand
    h = torch.relu(t.bn1(y @ t.W1 + t.b1))
    cr = C // cfg.num_branches
    z = torch.zeros(1, 9, 17, C)
    for r, d in enumerate(cfg.dilations):
        for j in range(3):
            for tt in range(9):
                src = tt + (j - 1) * d
                if 0 <= src < 9:
                    z[:, tt, :, r * cr:(r + 1) * cr] += h[:, src, :, r * cr:(r + 1) * cr] @ t.W2[r, j]
    ref = t.bn2(z + t.b2)
    assert torch.allclose(out, ref, atol=1e-5)


def test_se_gate_is_half_with_zero_fc():
    p = _params()
    for k in ("W1", "b1", "W2", "b2"):
        p[f"b0.se.{k}"] = np.zeros_like(p[f"b0.se.{k}"])
    se = segnet.ChannelAttention(p, "b0")
    u = torch.randn(2, 4, 17, 16)
    assert torch.equal(se.gate(u), torch.full((2, 16), 0.5))
    assert torch.equal(se(u), u * 0.5)


def test_stj_gate_shapes_and_range():
    p = _params()
    s = segnet.STJointAttention(p, "b0", 1e-5)
    x = torch.randn(2, 6, 17, 16)
    at, av = s.gates(x)
    assert at.shape == (2, 6, 16) and av.shape == (2, 17, 16)
    assert (at > 0).all() and (at < 1).all() and (av > 0).all() and (av < 1).all()
    assert torch.allclose(s(x), x * at[:, :, None] * av[:, None], atol=1e-7)


def test_folded_params_reproduce_unfolded_network():
    """BN folding (params.fold_params) == eval-mode BN in the oracle, <= 1e-5."""
    cfg = TINY
    p = _params()
    f = golfer_b200.params.fold_params(cfg, p)
    skel = torch.from_numpy(segnet.synth_skeletons(2, 10, cfg, seed=5))
    net = segnet.SegNet(cfg, p)
    with torch.no_grad():
        want = net.blocks[0].pre_attention(net.data_bn(skel.reshape(2, 10, -1)).reshape(skel.shape))
    V, Cin = cfg.num_joints, cfg.in_channels
    x = skel * torch.from_numpy(f["in.scale"]).reshape(V, Cin) + torch.from_numpy(f["in.shift"]).reshape(V, Cin)
    A = torch.from_numpy(f["b0.A"])
    xa = torch.einsum("pwv,btvc->btwpc", A, x).reshape(2, 10, V, 3 * Cin)
    y = torch.relu(xa @ torch.from_numpy(f["b0.Wg"]) + torch.from_numpy(f["b0.bg"]))
    h = torch.relu(y @ torch.from_numpy(f["b0.W1"]) + torch.from_numpy(f["b0.b1"]))
    C, R = 16, cfg.num_branches
    cr = C // R
    z = torch.zeros(2, 10, V, C)
    W2 = torch.from_numpy(f["b0.W2"])
    for r, d in enumerate(cfg.dilations):
        for j in range(3):
            for t in range(10):
                s = t + (j - 1) * d
                if 0 <= s < 10:
                    z[:, t, :, r * cr:(r + 1) * cr] += h[:, s, :, r * cr:(r + 1) * cr] @ W2[r, j]
    res = x @ torch.from_numpy(f["b0.Wr"]) + torch.from_numpy(f["b0.br"])
    got = torch.relu(z + torch.from_numpy(f["b0.b2"]) + res)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def test_blob_layout():
    cfg = TINY
    p = _params()
    blob = golfer_b200.params.pack_blob(cfg, p)
    head = blob[:4].view(np.uint32)
    assert head[0] == golfer_b200.params.BLOB_MAGIC
    assert head[1] == blob.size - 4 and head[2] == cfg.num_blocks
    f = golfer_b200.params.fold_params(cfg, p)
    assert sum(v.size for v in f.values()) == blob.size - 4
    assert np.array_equal(blob[4:4 + 51], f["in.scale"])


def test_golden_segnet(golden_dir):
    g = np.load(os.path.join(golden_dir, "segnet_small.npz"))
    for tag, cfg in (("tiny", TINY), ("v0", golfer_b200.V0)):
        params = golfer_b200.params.make_params(cfg, 1234)
        blob = golfer_b200.params.pack_blob(cfg, params)
        # the seeded generator is stable (numpy Generator stream + fp32 folding)
        assert golfer_b200.params.blob_sha256(blob) == str(g[f"{tag}_blob_sha256"])
        logits = segnet.segment_ref(cfg, params, g[f"{tag}_skel"])
        # torch CPU kernels may reorder sums across versions: pin to 1e-5, labels exactly
        np.testing.assert_allclose(logits, g[f"{tag}_logits"], rtol=1e-5, atol=1e-5)
        top2 = np.sort(g[f"{tag}_logits"], axis=-1)[..., -2:]
        safe = (top2[..., 1] - top2[..., 0]) > 1e-4
        assert np.array_equal(segnet.labels_from_logits(logits)[safe], g[f"{tag}_labels"][safe])


def test_label_rule_first_max():
    l = np.array([[[0.0, 1.0, 1.0], [2.0, 2.0, 1.0]]], np.float32)
    assert segnet.labels_from_logits(l).tolist() == [[1, 0]]


def test_adjacency_matches_independent_hop_distance_oracle():
    """the product's adjacency (params.build_adjacency, BFS + per-pair loop) against the
    oracle's independent formulation (matrix-power hop distances + masks, its own edge list)."""
    from oracle import adjacency
    A_product = golfer_b200.params.build_adjacency(golfer_b200.V0)
    A_oracle = adjacency.spatial_partitions()
    assert A_product.dtype == A_oracle.dtype == np.float32
    assert np.array_equal(A_product, A_oracle)
    # the edge lists were written independently: same undirected edge set
    mine = {tuple(sorted(e)) for limb in adjacency.EDGES_BY_LIMB.values() for e in limb}
    theirs = {tuple(sorted(e)) for e in golfer_b200.config.COCO_EDGES}
    assert mine == theirs and len(mine) == 16
    hop = adjacency.hop_distance_matrix()
    assert hop[0, 15] == 4 and hop[15, 16] == 8 and hop[3, 4] == 4     # ankle-nose, ankle-ankle, ear-ear
