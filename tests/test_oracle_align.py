"""Oracle self-consistency for the alignment path (CPU only; SURVEY.md section 4 tier 1)."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import align, align_native


def _rand_pair(Ta, Tb, seed, V=17, C=2):
    a, b = align.synth_swings(1, Ta, Tb, V=V, C=C, seed=seed)
    return a[0], b[0]


def test_self_alignment_is_diagonal_with_zero_cost():
    a, _ = _rand_pair(20, 20, 1)
    cost, path = align.align_ref(a, a)
    assert cost == 0.0
    assert np.array_equal(path, np.stack([np.arange(20), np.arange(20)], 1))


@pytest.mark.parametrize("Ta,Tb", [(1, 1), (1, 5), (5, 1), (4, 6), (6, 6), (6, 3)])
def test_dtw_matches_bruteforce(Ta, Tb):
    rng = np.random.default_rng(Ta * 10 + Tb)
    c = rng.uniform(0.1, 1.0, (Ta, Tb)).astype(np.float32)
    D, dirs = align.dtw_accumulate(c)
    assert np.isclose(D[-1, -1], align.dtw_bruteforce(c), rtol=1e-6)
    path = align.dtw_backtrack(dirs)
    assert tuple(path[0]) == (0, 0) and tuple(path[-1]) == (Ta - 1, Tb - 1)
    steps = np.diff(path, axis=0)
    assert set(map(tuple, steps)) <= {(1, 1), (1, 0), (0, 1)}
    assert max(Ta, Tb) <= len(path) <= Ta + Tb - 1
    # the path's own cost re-summed in path order equals D exactly (same add order)
    acc = np.float32(0)
    for i, j in path:
        acc = np.float32(c[i, j] + acc)
    assert acc == D[-1, -1]


def test_tie_break_prefers_diagonal_then_up_then_left():
    c = np.ones((3, 3), dtype=np.float32)
    _, dirs = align.dtw_accumulate(c)
    assert dirs[1, 1] == align.DIAG and dirs[2, 2] == align.DIAG
    # (2,1): diag D[1,0]=2, up D[1,1]=2, left D[2,0]=3  -> diagonal wins the tie
    assert dirs[2, 1] == align.DIAG
    c2 = np.array([[1, 5], [1, 1]], dtype=np.float32)
    # (1,1): diag=1, up=6, left=2 -> diag; make diag expensive to see up-vs-left tie
    c3 = np.array([[9, 1], [1, 1]], dtype=np.float32)
    _, d3 = align.dtw_accumulate(c3)   # diag 9, up 10, left 10 -> diag
    assert d3[1, 1] == align.DIAG
    c4 = np.array([[5, 0], [0, 1]], dtype=np.float32)  # diag 5, up 5, left 5 -> diag
    _, d4 = align.dtw_accumulate(c4)
    assert d4[1, 1] == align.DIAG
    c5 = np.array([[5, -1], [-1, 1]], dtype=np.float32)  # diag 5, up 4, left 4 -> up
    _, d5 = align.dtw_accumulate(c5)
    assert d5[1, 1] == align.UP
    assert align.dtw_accumulate(c2)[1][1, 1] == align.DIAG


def test_cost_symmetry_and_path_transpose_property():
    a, b = _rand_pair(14, 11, 5)
    cab = align.pair_cost(a, b)
    cba = align.pair_cost(b, a)
    assert np.array_equal(cab, cba.T)          # (-x)^2 == x^2 exactly
    Dab, _ = align.dtw_accumulate(cab)
    Dba, _ = align.dtw_accumulate(cba)
    assert np.array_equal(Dab, Dba.T)          # min is order-free, so D transposes exactly
    # paths transpose only up to the up/left tie preference; total cost is identical
    assert Dab[-1, -1] == Dba[-1, -1]


def test_pair_cost_known_answer():
    a = np.zeros((2, 17, 2), np.float32)
    b = np.zeros((3, 17, 2), np.float32)
    b[1, :, 0] = 3.0
    b[1, :, 1] = 4.0
    b[2, 0, 0] = 17.0
    c = align.pair_cost(a, b)
    assert np.array_equal(c[:, 0], [0, 0])
    assert np.array_equal(c[:, 1], [5, 5])
    assert np.array_equal(c[:, 2], [1, 1])


def test_extra_channels_are_ignored():
    a, b = _rand_pair(7, 9, 3, C=3)
    assert np.array_equal(align.pair_cost(a, b), align.pair_cost(a[..., :2], b[..., :2]))


def test_golden_align(golden_dir):
    g = np.load(os.path.join(golden_dir, "align_small.npz"))
    for tag in ("sq", "rect", "one"):
        a, b = g[f"{tag}_a"], g[f"{tag}_b"]
        for n in range(a.shape[0]):
            assert np.array_equal(align.pair_cost(a[n], b[n]), g[f"{tag}_cm"][n])
            cost, path = align.align_ref(a[n], b[n])
            L = g[f"{tag}_plen"][n]
            assert cost == g[f"{tag}_cost"][n]
            assert len(path) == L and np.array_equal(path, g[f"{tag}_path"][n, :L])
            assert np.all(g[f"{tag}_path"][n, L:] == -1)


def test_c_restatement_matches_numpy_bitwise(golden_dir):
    g = np.load(os.path.join(golden_dir, "align_small.npz"))
    for tag in ("sq", "rect", "one"):
        a, b = g[f"{tag}_a"], g[f"{tag}_b"]
        assert np.array_equal(align_native.pair_cost_c(a[0], b[0]), g[f"{tag}_cm"][0])
        for threads in (1, 4):
            cost, path, plen = align_native.align_batch_c(a, b, threads)
            assert np.array_equal(cost, g[f"{tag}_cost"])
            assert np.array_equal(plen, g[f"{tag}_plen"])
            assert np.array_equal(path, g[f"{tag}_path"])


@settings(max_examples=25, deadline=None)
@given(Ta=st.integers(1, 24), Tb=st.integers(1, 24), seed=st.integers(0, 2**16),
       cc=st.sampled_from([2, 3]))
def test_c_restatement_matches_numpy_property(Ta, Tb, seed, cc):
    a, b = align.synth_swings(2, Ta, Tb, C=cc, seed=seed)
    cost, path, plen = align_native.align_batch_c(a, b, 2)
    for n in range(2):
        c, p = align.align_ref(a[n], b[n])
        assert c == cost[n] and plen[n] == len(p)
        assert np.array_equal(path[n, :len(p)], p)


def test_c_restatement_full_size_pair():
    a, b = align.synth_swings(1, 300, 300, seed=7)
    cost, path, plen = align_native.align_batch_c(a, b, 1)
    c, p = align.align_ref(a[0], b[0])
    assert c == cost[0] and np.array_equal(path[0, :plen[0]], p)
    assert 300 <= plen[0] <= 599


def test_compare_ref_matches_cost_terms():
    a, b = _rand_pair(9, 8, 2)
    _, path = align.align_ref(a, b)
    d = align.compare_ref(a, b, path)
    assert d.shape == (len(path), 17)
    c = align.pair_cost(a, b)
    for l, (i, j) in enumerate(path):
        acc = np.float32(0)
        for v in range(17):
            acc = np.float32(acc + d[l, v])
        assert np.float32(acc / np.float32(17)) == c[i, j]


# ---- phase-conditioned alignment -------------------------------------------
def test_phase_penalty_zero_is_plain_alignment_and_c_matches_numpy():
    a, b = align.synth_swings(3, 23, 19, seed=5)
    rng = np.random.default_rng(0)
    la = rng.integers(0, 4, (3, 23)).astype(np.uint8)
    lb = rng.integers(0, 4, (3, 19)).astype(np.uint8)
    c0, p0, l0 = align_native.align_batch_c(a, b)
    c1, p1, l1 = align_native.align_phase_batch_c(a, b, la, lb, 0.0)
    assert np.array_equal(c0, c1) and np.array_equal(p0, p1) and np.array_equal(l0, l1)
    for pen in (0.25, 3.0, np.inf):
        cc, pc, lc = align_native.align_phase_batch_c(a, b, la, lb, pen)
        for n in range(3):
            cost, path = align.align_phase_ref(a[n], b[n], la[n], lb[n], pen)
            assert cost == cc[n] or (np.isinf(cost) and np.isinf(cc[n]))
            assert lc[n] == len(path) and np.array_equal(pc[n, :len(path)], path)


def test_phase_penalty_keeps_the_path_inside_matching_phases():
    # two clips with the same three phases but different durations: with a large penalty the path
    # only visits cells whose labels agree (such a path exists: the phase order is the same)
    la = np.repeat(np.uint8([0, 1, 2]), [6, 10, 4])
    lb = np.repeat(np.uint8([0, 1, 2]), [9, 3, 8])
    a, _ = align.synth_swings(1, 20, 20, seed=11)
    _, b = align.synth_swings(1, 20, 20, seed=12)
    cost, path = align.align_phase_ref(a[0], b[0], la, lb, 1e6)
    assert np.all(la[path[:, 0]] == lb[path[:, 1]])
    assert cost < 1e6
    plain_cost, plain_path = align.align_ref(a[0], b[0])
    assert cost >= plain_cost
    # the unconstrained optimum crosses phases here, so the constraint really binds
    assert np.any(la[plain_path[:, 0]] != lb[plain_path[:, 1]])


def test_division_by_17_in_three_operations_is_the_ieee_quotient(tmp_path):
    """The DTW cost producers divide the joint sum by V = 17 as q0 = x * rc, r = fma(-17, q0, x), q = fma(r, rc, q0)
    (csrc/align.cu:div17_exact) instead of a full division.  tools/div17_check.c compares that with x / 17 for
    every normal float32 (45 s); this runs every 211th value of both signs (20 M values)."""
    import subprocess
    exe = tmp_path / "div17_check"
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "div17_check.c")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), src, "-lm"], check=True)
    out = subprocess.run([str(exe), "211"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0" in out.stdout
