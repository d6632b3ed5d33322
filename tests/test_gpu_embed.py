"""GPU parity of the learned alignment embedding against oracle/embed.py, under the policy
its header declares: cost matrix within 1e-2 (fp32 oracle) / 1e-3 (bf16-emulation oracle) relative; DTW total and
path BIT-EXACT against the oracle's DP run on the GPU's own cost matrix; total within 1e-2 of the fp32 oracle;
path agreement with the fp32 oracle reported."""
import numpy as np
import pytest
import torch

import golfer_b200
from oracle import align as oalign
from oracle import embed as oembed

pytestmark = pytest.mark.gpu


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def aligner():
    p = oembed.make_embed_params()
    return p, golfer_b200.EmbedAligner(oembed.pack_embed_blob(p))


@pytest.mark.parametrize("N,Ta,Tb,Cc", [(3, 300, 300, 2), (3, 300, 257, 3), (3, 257, 300, 2), (4, 64, 40, 2), (4, 40, 64, 2),
                                        (2, 1, 1, 2), (3, 1, 33, 2), (3, 33, 1, 2), (5, 129, 128, 2), (2, 500, 420, 2),
                                        (300, 24, 24, 2)])
def test_embed_alignment_against_oracle(aligner, N, Ta, Tb, Cc):
    p, al = aligner
    a, b = oalign.synth_swings(N, Ta, Tb, C=Cc, seed=Ta * 3 + Tb)
    cost, path, plen, cm = al.align(_dev(a), _dev(b), want_cost_matrix=True)
    cost, path, plen, cm = cost.cpu().numpy(), path.cpu().numpy(), plen.cpu().numpy(), cm.cpu().numpy()
    agree = []
    for n in range(min(N, 6)):
        want32 = oembed.embed_cost(a[n], b[n], p)
        want16 = oembed.embed_cost(a[n], b[n], p, emulate_bf16=True)
        got = cm[n] if Ta >= Tb else cm[n].T                       # the matrix is stored for the exchanged pair when Ta < Tb
        scale = max(float(want32.max()), 1e-6)
        assert np.abs(got - want32).max() / scale < 1e-2, n
        assert np.abs(got - want16).max() / scale < 1e-3, n
        # the DP and backtrack kernels, pinned on the GPU's own cost matrix
        total, ref_path = oembed.dtw_on_cost(got)
        assert cost[n] == total, (n, cost[n], total)
        assert plen[n] == len(ref_path) and np.array_equal(path[n, :plen[n]], ref_path), n
        assert np.all(path[n, plen[n]:] == -1)
        t32, p32, _ = oembed.align_embed_ref(a[n], b[n], p)
        assert abs(cost[n] - t32) <= 1e-2 * max(abs(t32), 1e-6)
        agree.append(float(np.mean([tuple(c) in set(map(tuple, p32)) for c in ref_path])))
    print(f"N={N} {Ta}x{Tb}: path cells shared with the fp32 oracle's path: {np.mean(agree):.3f}")


def test_embed_cost_only_and_errors(aligner):
    p, al = aligner
    a, b = oalign.synth_swings(10, 50, 44, seed=3)
    c1, p1, l1 = al.align(_dev(a), _dev(b))
    c2, p2, l2 = al.align(_dev(a), _dev(b), want_path=False)
    assert p2 is None and torch.equal(c1, c2)
    with pytest.raises(golfer_b200.GolferError):
        al.align(_dev(a), _dev(b[:5]))
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.EmbedAligner(np.zeros(100, np.float32))
    a25, b25 = oalign.synth_swings(2, 10, 10, V=25, seed=1)
    with pytest.raises(golfer_b200.GolferError):
        al.align(_dev(a25), _dev(b25))
