"""Per-kernel parity on the bf16 path ("kernel parity"): every kernel of a block is
checked in ISOLATION, i.e. against the oracle sub-function applied to the kernel's OWN input as the GPU
produced it (read back through gs_debug_read), so an error cannot hide behind, or be blamed on, its
neighbours.  v0 widths, T in {1, 7, 8, 129, 300} (one frame, one GCN tile, tile + 1, temporal tile + 9, headline).

  front_mma_kernel / gcn_fused_kernel   XA, Y   vs  oracle GraphConv            (README.md:27-28)
  tcn_fused_kernel                      U       vs  MultiBranchTCN + residual   (README.md:29-30)
  tcn epilogue pooling                  PT, PV  vs  sums of U
  se_kernel                             seS     vs  ChannelAttention.gate       (README.md:31-32)
  stj_tc_kernel                         gT, gV  vs  STJointAttention.gates      (README.md:33-34)
Tolerances are relative to max|want| of each tensor: 1e-2 for tensors that carry bf16 roundings of
operands / outputs, 2e-3 for the fp32 gate math (fp16 mma.sync inputs in the ST-joint kernel).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

os.environ["GOLFER_DEBUG_XA"] = "1"          # the GCN kernel also dumps its XA chunks (read at context creation)

import golfer_b200                           # noqa: E402
from oracle import segnet as osegnet         # noqa: E402

pytestmark = pytest.mark.gpu
CFG = golfer_b200.V0
V = 17


def _rel(got, want):
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-20))


def _bf16(ctx, name, shape):
    n = int(np.prod(shape))
    raw = ctx.debug_read(name, n * 2)
    return torch.from_numpy(raw.copy()).view(torch.bfloat16).reshape(shape).float()


def _f32(ctx, name, shape):
    n = int(np.prod(shape))
    raw = ctx.debug_read(name, n * 4)
    return torch.from_numpy(raw.copy()).view(torch.float32).reshape(shape).clone()


@pytest.fixture(scope="module")
def net():
    params = golfer_b200.params.make_params(CFG, 1234)
    return params, osegnet.SegNet(CFG, params)


@pytest.mark.parametrize("T", [1, 7, 8, 129, 300])
def test_each_kernel_against_its_oracle_stage(net, T):
    params, onet = net
    B = 2
    skel = osegnet.synth_skeletons(B, T, CFG, seed=100 + T)
    seg = golfer_b200.Segmenter(CFG, params, precision="bf16", max_B=B, max_T=T)
    x = torch.from_numpy(skel).cuda()
    worst = {}
    with torch.no_grad():
        for i, (cin, c) in enumerate(CFG.block_io()):
            seg.features(x, i)                      # runs blocks 0..i; the workspace now holds block i's tensors
            ctx, blk = seg.ctx, onet.blocks[i]
            Y = _bf16(ctx, "Y", (B, V, T, c)).permute(0, 2, 1, 3).contiguous()          # joint-major -> [B,T,V,C]
            U = _bf16(ctx, f"U{i & 1}", (B, T, V, c))
            if i == 0:
                xin = onet.data_bn(torch.from_numpy(skel).reshape(B, T, V * cin)).reshape(B, T, V, cin)
                res = _bf16(ctx, "R", (B, T, V, c))
                worst[f"b{i}.R"] = _rel(res, blk.bnr(xin @ blk.Wr + blk.br))
            else:
                xin = _bf16(ctx, "X", (B, T, V, cin))                                     # the gated input as stored
                res = blk.bnr(xin @ blk.Wr + blk.br) if blk.has_res else xin
                # XA dump: [tile][128 rows (7 frames x 17 joints, then padding)][3*Cin], column p*Cin + c
                ntile = (T + 6) // 7
                xa = _bf16(ctx, "XA", (B * ntile, 128, 3 * cin))[:, :119].reshape(B, ntile * 7, V, 3, cin)[:, :T]
                want_xa = torch.einsum("pwv,btvc->btwpc", blk.gcn.A, xin)
                worst[f"b{i}.XA"] = _rel(xa, want_xa)
            worst[f"b{i}.Y"] = _rel(Y, blk.gcn(xin))
            worst[f"b{i}.U"] = _rel(U, F.relu(blk.tcn(Y) + res))
            # pooling sums taken in the temporal kernel's epilogue (fp32 values before the bf16 rounding of U)
            worst[f"b{i}.PT"] = _rel(_f32(ctx, "PT", (B, T, c)), U.sum(2))
            worst[f"b{i}.PV"] = _rel(_f32(ctx, "PV", (B, V, c)), U.sum(1))
            se = _f32(ctx, "seS", (B, c))
            worst[f"b{i}.se"] = _rel(se, blk.se.gate(U))
            at, av = blk.stj.gates(U * se[:, None, None, :])
            worst[f"b{i}.gT"] = _rel(_f32(ctx, "gT", (B, T, c)), se[:, None, :] * at)
            worst[f"b{i}.gV"] = _rel(_f32(ctx, "gV", (B, V, c)), av)
    print(f"T={T}: " + "  ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    for k, v in worst.items():
        tol = 2e-3 if k.split(".")[1] in ("se", "gT", "gV") else 1e-2
        assert v < tol, (k, v)
    seg.ctx.close()
