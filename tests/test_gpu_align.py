"""GPU parity for the alignment path, through the C ABI (bit-exact bar)."""
import os

import numpy as np
import pytest
import torch

import golfer_b200
from oracle import align as oalign
from oracle import align_native

pytestmark = pytest.mark.gpu


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _check_against(a, b, cost, path, plen, ref_cost, ref_path, ref_plen):
    assert np.array_equal(cost.cpu().numpy(), ref_cost)
    assert np.array_equal(plen.cpu().numpy(), ref_plen)
    assert np.array_equal(path.cpu().numpy(), ref_path)


def test_golden_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "align_small.npz"))
    for tag in ("sq", "rect", "one"):
        a, b = g[f"{tag}_a"], g[f"{tag}_b"]
        cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
        _check_against(a, b, cost, path, plen, g[f"{tag}_cost"], g[f"{tag}_path"], g[f"{tag}_plen"])
        cm = golfer_b200.host.pair_cost(_dev(a), _dev(b))
        assert np.array_equal(cm.cpu().numpy(), g[f"{tag}_cm"])


@pytest.mark.parametrize("N,Ta,Tb,Cc", [(5, 300, 300, 2), (4, 300, 257, 3), (3, 64, 300, 2), (2, 1, 1, 2),
                                        (3, 1, 33, 2), (3, 33, 1, 2), (6, 17, 31, 2), (2, 500, 420, 2),
                                        # many more pairs than resident CTAs (each CTA pipelines several pairs),
                                        # tiny shapes, odd / even column counts around the two-column mapping
                                        (9000, 24, 24, 2), (7001, 9, 5, 2), (4, 2, 2, 2), (5, 3, 2, 2), (3, 65, 64, 2),
                                        (3, 66, 65, 3), (2, 1024, 1000, 2)])
def test_matches_oracle_bitwise(N, Ta, Tb, Cc):
    a, b = oalign.synth_swings(N, Ta, Tb, C=Cc, seed=Ta * 7 + Tb)
    ref_cost, ref_path, ref_plen = align_native.align_batch_c(a, b, 4)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
    _check_against(a, b, cost, path, plen, ref_cost, ref_path, ref_plen)


def test_generic_kernel_path_large_and_odd_joint_count():
    # Tb > 1024 and V != 17 both route to the generic (global-scratch) kernel
    a, b = oalign.synth_swings(2, 40, 1100, seed=3)
    rc, rp, rl = align_native.align_batch_c(a, b, 4)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
    _check_against(a, b, cost, path, plen, rc, rp, rl)
    a, b = oalign.synth_swings(3, 50, 45, V=25, seed=4)
    rc, rp, rl = align_native.align_batch_c(a, b, 4)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
    _check_against(a, b, cost, path, plen, rc, rp, rl)


def test_sqrt_range_edges_bitwise():
    # The wavefront kernel takes square roots with a packed refinement that is exact only for
    # 2^-101 <= x < inf and falls back per frame otherwise: exercise both sides of every edge.
    rng = np.random.default_rng(11)
    a, b = oalign.synth_swings(8, 60, 70, seed=21)
    b[0, :, 3] = a[0, :60, 3].mean(0)                 # pair 0: ordinary
    b[1, :60] = a[1]                                  # pair 1: coincident joints on the diagonal (x = 0)
    a[2] *= np.float32(2.0 ** -52); b[2] *= np.float32(2.0 ** -52)     # x around 2^-104: below the range
    a[3] *= np.float32(2.0 ** -50); b[3] *= np.float32(2.0 ** -50)     # x straddling 2^-101
    a[4] *= np.float32(2.0 ** -70); b[4] *= np.float32(2.0 ** -70)     # squares underflow to denormals / 0
    a[5] *= np.float32(2.0 ** 62); b[5] *= np.float32(2.0 ** 62)       # x near 2^126: still finite
    a[6] *= np.float32(2.0 ** 64); b[6] *= np.float32(2.0 ** 64)       # some squares overflow to inf
    a[7, rng.integers(0, 60, 5), rng.integers(0, 17, 5), 0] = 0.0      # exact zeros against non-zeros
    b[7, 10, 4] = a[7, 10, 4]                                           # one coincident joint in one frame
    with np.errstate(over="ignore", invalid="ignore", under="ignore"):
        ref_cost, ref_path, ref_plen = align_native.align_batch_c(a, b, 4)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
    assert np.array_equal(cost.cpu().numpy(), ref_cost, equal_nan=True)
    assert np.array_equal(plen.cpu().numpy(), ref_plen)
    assert np.array_equal(path.cpu().numpy(), ref_path)
    cm = golfer_b200.host.pair_cost(_dev(a), _dev(b)).cpu().numpy()
    with np.errstate(over="ignore", invalid="ignore", under="ignore"):
        for n in range(8):
            assert np.array_equal(cm[n], oalign.pair_cost(a[n], b[n]), equal_nan=True), n


def test_self_alignment_zero_cost_diagonal():
    a, _ = oalign.synth_swings(2, 300, 300, seed=9)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(a))
    assert torch.all(cost == 0)
    assert torch.all(plen == 300)
    diag = torch.arange(300, dtype=torch.int32, device="cuda")
    assert torch.equal(path[0, :300, 0], diag) and torch.equal(path[0, :300, 1], diag)
    assert torch.all(path[:, 300:] == -1)


def test_cost_only_mode_and_host_entry_point():
    a, b = oalign.synth_swings(70, 120, 100, seed=21)
    rc, rp, rl = align_native.align_batch_c(a, b, 4)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b), want_path=False)
    assert path is None and np.array_equal(cost.cpu().numpy(), rc)
    # host buffers in, host buffers out (H2D / D2H inside gs_align_host, chunked)
    cost_h, path_h, plen_h = golfer_b200.host.align_batch(a, b)
    assert not cost_h.is_cuda
    assert np.array_equal(cost_h.numpy(), rc) and np.array_equal(path_h.numpy(), rp)
    assert np.array_equal(plen_h.numpy(), rl)


def test_host_entry_point_many_chunks():
    """gs_align_host cuts a large batch into chunks of whole sweep rounds (a multiple of the SM count, ~16 chunks):
    700 pairs = 4 chunks of 148 + 108 on a 148-SM part; results must not depend on the chunking."""
    a, b = oalign.synth_swings(700, 24, 20, seed=5)
    rc, rp, rl = align_native.align_batch_c(a, b, 4)
    cost_h, path_h, plen_h = golfer_b200.host.align_batch(a, b)
    assert np.array_equal(cost_h.numpy(), rc) and np.array_equal(path_h.numpy(), rp)
    assert np.array_equal(plen_h.numpy(), rl)
    cost_d, path_d, plen_d = golfer_b200.host.align_batch(_dev(a), _dev(b))
    assert np.array_equal(cost_d.cpu().numpy(), rc) and np.array_equal(path_d.cpu().numpy(), rp)


def test_pipelined_host_entry_point_matches_oracle():
    """align_submit / align_wait: two batches in flight on two sets of staging buffers, other entry points in between."""
    ctx = golfer_b200.host.Context(0)
    shapes = [(300, 24, 20), (64, 33, 40), (310, 24, 20), (5, 17, 9), (300, 24, 20)]
    data = []
    for i, (N, Ta, Tb) in enumerate(shapes):
        a, b = oalign.synth_swings(N, Ta, Tb, seed=60 + i)
        data.append((torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory(),
                     align_native.align_batch_c(a, b, 4)))
    tks = [golfer_b200.host.align_submit(a, b, ctx=ctx) for a, b, _ in data[:3]]     # the third waits for the first inside
    dev_cost, dev_path, _ = golfer_b200.host.align_batch(data[3][0].cuda(), data[3][1].cuda(), ctx=ctx)
    tks.append(golfer_b200.host.align_submit(data[4][0], data[4][1], ctx=ctx, want_path=False))
    for tk, (_, _, (rc, rp, rl)) in zip(tks[:3], data[:3]):
        cost, path, plen = golfer_b200.host.align_wait(tk)
        assert np.array_equal(cost.numpy(), rc) and np.array_equal(path.numpy(), rp) and np.array_equal(plen.numpy(), rl)
    assert np.array_equal(dev_cost.cpu().numpy(), data[3][2][0]) and np.array_equal(dev_path.cpu().numpy(), data[3][2][1])
    cost, path, plen = golfer_b200.host.align_wait(tks[3])
    assert path is None and np.array_equal(cost.numpy(), data[4][2][0])
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.host.align_submit(data[0][0].cuda(), data[0][1], ctx=ctx)
    ctx.close()


def test_public_align_signature_single_pair():
    a, b = oalign.synth_swings(1, 30, 26, seed=2)
    cost, path = golfer_b200.align(_dev(a[0]), _dev(b[0]))
    c, p = oalign.align_ref(a[0], b[0])
    assert cost.item() == c and np.array_equal(path.cpu().numpy(), p)


def test_full_size_batch_properties_and_sampled_parity():
    """BASELINE.json configs[2]: 4096 pairs of 300x300.  Whole batch vs the C oracle
    (a few seconds on the host), plus size-independent path properties."""
    N = 4096
    a, b = oalign.synth_swings(N, 300, 300, seed=7)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
    torch.cuda.synchronize()
    rc, rp, rl = align_native.align_batch_c(a, b, os.cpu_count() or 1)
    assert np.array_equal(cost.cpu().numpy(), rc)
    assert np.array_equal(plen.cpu().numpy(), rl)
    assert np.array_equal(path.cpu().numpy(), rp)
    p = path.cpu().numpy()
    L = plen.cpu().numpy()
    assert np.all(p[:, 0] == 0) and np.all(L >= 300) and np.all(L <= 599)
    for n in range(0, N, 257):
        steps = np.diff(p[n, :L[n]], axis=0)
        assert set(map(tuple, steps)) <= {(1, 1), (1, 0), (0, 1)}
        assert tuple(p[n, L[n] - 1]) == (299, 299)


def test_compare_two_skeletons():
    a, b = oalign.synth_swings(3, 40, 36, seed=5)
    da, db = _dev(a), _dev(b)
    cost, path, plen = golfer_b200.host.align_batch(da, db)
    out = golfer_b200.compare(da, db, path, plen).cpu().numpy()
    for n in range(3):
        L = int(plen[n])
        ref = oalign.compare_ref(a[n], b[n], path[n, :L].cpu().numpy())
        assert np.array_equal(out[n, :L], ref)
        assert np.all(out[n, L:] == 0)


def test_bad_arguments_raise():
    a = torch.zeros(2, 8, 17, 2, device="cuda")
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.host.align_batch(a, torch.zeros(3, 8, 17, 2, device="cuda"))
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.host.align_batch(a[..., :1], a[..., :1])


# ---- phase-conditioned alignment ------------------------------------------------
def _phase_labels(N, T, seed, nphase=5):
    """Monotone phase labels with random run lengths (what a segmentation of a swing looks like)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((N, T), np.uint8)
    for n in range(N):
        cuts = np.sort(rng.integers(0, T + 1, nphase - 1))
        out[n] = np.searchsorted(cuts, np.arange(T), side="right")
    return out


@pytest.mark.parametrize("N,Ta,Tb,pen", [(6, 300, 300, 0.5), (5, 300, 257, 2.0), (4, 64, 300, 0.5), (3, 1, 1, 1.0),
                                         (4, 33, 1, 1.0), (5, 300, 300, float("inf")), (4, 120, 90, 0.0),
                                         (700, 40, 40, 0.75), (9000, 20, 19, 0.3), (3, 2, 1, 1.5), (4, 65, 33, 0.25)])
def test_phase_alignment_matches_oracle_bitwise(N, Ta, Tb, pen):
    a, b = oalign.synth_swings(N, Ta, Tb, seed=Ta * 5 + Tb)
    la, lb = _phase_labels(N, Ta, 1), _phase_labels(N, Tb, 2)
    ref_cost, ref_path, ref_plen = align_native.align_phase_batch_c(a, b, la, lb, pen, 4)
    cost, path, plen = golfer_b200.align_phase(_dev(a), _dev(b), _dev(la), _dev(lb), pen)
    assert np.array_equal(cost.cpu().numpy(), ref_cost)
    assert np.array_equal(plen.cpu().numpy(), ref_plen)
    assert np.array_equal(path.cpu().numpy(), ref_path)
    cost_only, _, _ = golfer_b200.align_phase(_dev(a), _dev(b), _dev(la), _dev(lb), pen, want_path=False)
    assert np.array_equal(cost_only.cpu().numpy(), ref_cost)


def test_segment_labels_feed_phase_alignment_on_device():
    # segment -> labels -> align_phase without a host trip; checked against the oracle run on the
    # labels the GPU produced (label parity itself is tests/test_gpu_segment.py's subject)
    cfg = golfer_b200.V0
    seg = golfer_b200.Segmenter(cfg, seed=1234, precision="bf16", max_B=8, max_T=96)
    from oracle import segnet as osegnet
    skel = torch.from_numpy(osegnet.synth_skeletons(8, 96, cfg, seed=4)).cuda()
    _, labels = seg.segment(skel, return_labels=True)
    a, b = skel[0::2, :, :, :2].contiguous(), skel[1::2, :, :, :2].contiguous()
    la, lb = labels[0::2].contiguous(), labels[1::2].contiguous()
    cost, path, plen = golfer_b200.align_phase(a, b, la, lb, 0.5)
    rc, rp, rl = align_native.align_phase_batch_c(a.cpu().numpy(), b.cpu().numpy(), la.cpu().numpy(),
                                                  lb.cpu().numpy(), 0.5, 2)
    assert np.array_equal(cost.cpu().numpy(), rc) and np.array_equal(path.cpu().numpy(), rp)
    assert np.array_equal(plen.cpu().numpy(), rl)
    seg.ctx.close()


def test_phase_alignment_rejects_bad_arguments():
    a, b = oalign.synth_swings(2, 10, 10, seed=1)
    la = np.zeros((2, 10), np.uint8)
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.align_phase(_dev(a), _dev(b), _dev(la), _dev(la[:, :5]), 1.0)
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.align_phase(a, b, la, la, 1.0)                      # host arrays
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.align_phase(_dev(a), _dev(b), _dev(la), _dev(la), float("nan"))


def test_four_byte_aligned_inputs():
    # a view whose data pointer is only 4-byte aligned: the staging copies must fall back from 8-byte cp.async
    N, Ta, Tb = 5, 90, 70
    a, b = oalign.synth_swings(N, Ta, Tb, seed=77)
    fa = torch.zeros(a.size + 1, device="cuda")
    fb = torch.zeros(b.size + 1, device="cuda")
    fa[1:] = torch.from_numpy(a).cuda().flatten()
    fb[1:] = torch.from_numpy(b).cuda().flatten()
    va, vb = fa[1:].view(N, Ta, 17, 2), fb[1:].view(N, Tb, 17, 2)
    assert va.data_ptr() % 8 == 4 and va.is_contiguous()
    ref_cost, ref_path, ref_plen = align_native.align_batch_c(a, b, 4)
    cost, path, plen = golfer_b200.host.align_batch(va, vb)
    _check_against(a, b, cost, path, plen, ref_cost, ref_path, ref_plen)


# ---- non-finite values on the boundary (row 0 steps LEFT, column 0 steps UP whatever D holds) ----------
@pytest.mark.parametrize("N,Ta,Tb,V", [(5, 6, 4, 17), (5, 4, 6, 17), (4, 70, 41, 17), (4, 41, 70, 17), (3, 300, 257, 17),
                                       (3, 257, 300, 17), (3, 9, 5, 25), (3, 5, 9, 25), (2, 1, 7, 17), (2, 7, 1, 17)])
def test_infinite_penalty_with_mismatched_first_labels(N, Ta, Tb, V):
    """penalty = +inf and la[:,0] != lb[:,0]: D[0,0] = inf, so every D is inf and every comparison of the DP
    step is false; the path must still be the oracle's (forced LEFT on row 0, UP on column 0, DIAG inside)."""
    a, b = oalign.synth_swings(N, Ta, Tb, V=V, seed=Ta * 3 + Tb)
    la = np.zeros((N, Ta), np.uint8)
    lb = np.ones((N, Tb), np.uint8)
    la[0::2, Ta // 2:] = 1                       # some pairs agree further in; the first frames never do
    ref_cost, ref_path, ref_plen = align_native.align_phase_batch_c(a, b, la, lb, float("inf"), 2)
    cost, path, plen = golfer_b200.align_phase(_dev(a), _dev(b), _dev(la), _dev(lb), float("inf"))
    assert np.array_equal(cost.cpu().numpy(), ref_cost)
    assert np.array_equal(plen.cpu().numpy(), ref_plen)
    assert np.array_equal(path.cpu().numpy(), ref_path)


@pytest.mark.parametrize("N,Ta,Tb,V", [(6, 6, 4, 17), (6, 4, 6, 17), (6, 90, 61, 17), (6, 61, 90, 17), (6, 12, 7, 25),
                                       (6, 7, 12, 25), (6, 40, 1100, 17)])
def test_nan_and_inf_keypoints_rectangular(N, Ta, Tb, V):
    """NaN / inf keypoints on and off the boundary with Ta != Tb, on the pipelined kernel (both orientations)
    and the generic kernel (V != 17, more than 1024 columns)."""
    a, b = oalign.synth_swings(N, Ta, Tb, V=V, seed=Ta * 11 + Tb)
    a[0, 0, 0, 0] = np.nan                       # first student frame: the whole first row is NaN
    b[1, 0, 2, 1] = np.inf                       # first reference frame: the whole first column is inf
    a[2, Ta - 1, 1, 0] = np.nan                  # last row
    b[3, Tb // 2, 0, 0] = -np.inf                # an interior column
    a[4, :, 3, 1] = np.nan                       # everything NaN
    with np.errstate(over="ignore", invalid="ignore"):
        ref_cost, ref_path, ref_plen = align_native.align_batch_c(a, b, 2)
    cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
    torch.cuda.synchronize()
    assert np.array_equal(cost.cpu().numpy(), ref_cost, equal_nan=True)
    assert np.array_equal(plen.cpu().numpy(), ref_plen)
    assert np.array_equal(path.cpu().numpy(), ref_path)


@pytest.mark.parametrize("N,Ta,Tb,pen", [(5, 257, 300, 2.0), (4, 1, 33, 1.0), (700, 31, 40, 0.75), (3, 1, 2, 1.5),
                                         (4, 33, 65, 0.25), (2, 1000, 1024, 0.5)])
def test_phase_alignment_shorter_student(N, Ta, Tb, pen):
    """Ta < Tb with phase labels: same pipelined kernel, sequences exchanged inside the launch."""
    a, b = oalign.synth_swings(N, Ta, Tb, seed=Ta * 5 + Tb)
    la, lb = _phase_labels(N, Ta, 3), _phase_labels(N, Tb, 4)
    ref_cost, ref_path, ref_plen = align_native.align_phase_batch_c(a, b, la, lb, pen, 4)
    cost, path, plen = golfer_b200.align_phase(_dev(a), _dev(b), _dev(la), _dev(lb), pen)
    assert np.array_equal(cost.cpu().numpy(), ref_cost)
    assert np.array_equal(plen.cpu().numpy(), ref_plen)
    assert np.array_equal(path.cpu().numpy(), ref_path)


def test_ties_resolve_identically_in_both_orientations():
    """Quantised keypoints make exact ties between up / left / diagonal common: the exchanged-sequence launch
    (Ta < Tb) must break them exactly as the oracle does in the original orientation."""
    rng = np.random.default_rng(5)
    for Ta, Tb in ((40, 64), (64, 40), (33, 34), (34, 33)):
        a = rng.integers(0, 3, (16, Ta, 17, 2)).astype(np.float32)
        b = rng.integers(0, 3, (16, Tb, 17, 2)).astype(np.float32)
        ref_cost, ref_path, ref_plen = align_native.align_batch_c(a, b, 2)
        cost, path, plen = golfer_b200.host.align_batch(_dev(a), _dev(b))
        _check_against(a, b, cost, path, plen, ref_cost, ref_path, ref_plen)


def test_device_call_then_host_call_share_workspace_safely():
    """A device-stream call followed at once by a host-buffer call on the same context (its own streams):
    the second must wait for the first (ctx ordering event), results equal the oracle's."""
    a, b = oalign.synth_swings(600, 120, 100, seed=31)
    a2, b2 = oalign.synth_swings(600, 120, 100, seed=32)
    rc1, rp1, _ = align_native.align_batch_c(a, b, 4)
    rc2, rp2, _ = align_native.align_batch_c(a2, b2, 4)
    ctx = golfer_b200.host.Context(0)
    side = torch.cuda.Stream()
    da, db = _dev(a), _dev(b)
    torch.cuda.synchronize()
    for _ in range(3):
        with torch.cuda.stream(side):
            c1, p1, _ = golfer_b200.host.align_batch(da, db, ctx=ctx)
        c2, p2, _ = golfer_b200.host.align_batch(a2, b2, ctx=ctx)          # host entry point, own streams
        torch.cuda.synchronize()
        assert np.array_equal(c1.cpu().numpy(), rc1) and np.array_equal(p1.cpu().numpy(), rp1)
        assert np.array_equal(c2.numpy(), rc2) and np.array_equal(p2.numpy(), rp2)
    ctx.close()


def test_compare_validates_devices_dtypes_and_path_range():
    a, b = oalign.synth_swings(2, 12, 9, seed=6)
    da, db = _dev(a), _dev(b)
    cost, path, plen = golfer_b200.host.align_batch(da, db)
    want = golfer_b200.compare(da, db, path, plen)
    got = golfer_b200.compare(da, db, path.to(torch.int64), plen.to(torch.int64))   # converted, not reinterpreted
    assert torch.equal(got, want)
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.compare(da, torch.from_numpy(b), path, plen)                    # host tensor
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.host.align_batch(da, torch.from_numpy(b))
    bad = path.clone()
    bad[0, 0, 0] = 12                                                                # past the end of a
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.compare(da, db, bad, plen)
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.compare(da, db, path, plen + 100)
