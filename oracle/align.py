"""CPU ORACLE (test infrastructure, never shipped or timed as the product):
NumPy fp32 restatement of the temporal-alignment step.

PARITY UNPINNED: the reference ships no code, tests or golden vectors
(SURVEY.md section 0 / 8c).  This module is the DEFINITION of correct for
`align`.  Reference evidence: README.md:21-22 ("temporal alignment model"),
README.md:44-49 (alignment section), README.md:50-52 ("Compare 2 skeleton").

Arithmetic contract (what the CUDA kernels must reproduce bit-for-bit):
  cost[i,j] = (sum_{v=0..V-1, sequential} sqrt(dx*dx + dy*dy)) / V
      every operation an individually rounded IEEE fp32 op: NO fused multiply-add,
      correctly rounded sqrt and division, joints summed in index order.
  D[0,0] = c[0,0];  D[i,0] = c[i,0] + D[i-1,0];  D[0,j] = c[0,j] + D[0,j-1]
  D[i,j] = c[i,j] + min(D[i-1,j-1], D[i-1,j], D[i,j-1])
  step taken: first minimum in the order diagonal, up (i-1,j), left (i,j-1).
  path = cells from (0,0) to (Ta-1,Tb-1); cost = D[Ta-1,Tb-1] (un-normalised).
"""
from __future__ import annotations

import numpy as np

DIAG, UP, LEFT = 0, 1, 2


def pair_cost(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a [Ta,V,>=2], b [Tb,V,>=2] fp32 -> cost [Ta,Tb] fp32.

    Mean over joints of the 2-D Euclidean distance; explicit v = 0..V-1 loop so
    the summation order is sequential (np.sum would be pairwise)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    Ta, V = a.shape[0], a.shape[1]
    Tb = b.shape[0]
    acc = np.zeros((Ta, Tb), dtype=np.float32)
    for v in range(V):
        dx = a[:, v, 0][:, None] - b[:, v, 0][None, :]
        dy = a[:, v, 1][:, None] - b[:, v, 1][None, :]
        s = dx * dx
        s = s + dy * dy                 # two rounded products, one rounded sum: no FMA
        acc = acc + np.sqrt(s)
    return acc / np.float32(V)


def dtw_accumulate(c: np.ndarray):
    """cost [Ta,Tb] fp32 -> (D [Ta,Tb] fp32, dirs [Ta,Tb] u8)  (SURVEY 8a row a8)."""
    c = np.asarray(c, dtype=np.float32)
    Ta, Tb = c.shape
    D = np.empty((Ta, Tb), dtype=np.float32)
    dirs = np.zeros((Ta, Tb), dtype=np.uint8)
    D[0, 0] = c[0, 0]
    for j in range(1, Tb):
        D[0, j] = c[0, j] + D[0, j - 1]
        dirs[0, j] = LEFT
    for i in range(1, Ta):
        D[i, 0] = c[i, 0] + D[i - 1, 0]
        dirs[i, 0] = UP
        Dp, Di, ci = D[i - 1], D[i], c[i]
        for j in range(1, Tb):
            best, d = Dp[j - 1], DIAG
            if Dp[j] < best:
                best, d = Dp[j], UP
            if Di[j - 1] < best:
                best, d = Di[j - 1], LEFT
            Di[j] = ci[j] + best
            dirs[i, j] = d
    return D, dirs


def dtw_backtrack(dirs: np.ndarray) -> np.ndarray:
    """dirs [Ta,Tb] -> path [L,2] int32 from (0,0) to (Ta-1,Tb-1)  (SURVEY 8a row a9)."""
    Ta, Tb = dirs.shape
    i, j = Ta - 1, Tb - 1
    rev = [(i, j)]
    while i > 0 or j > 0:
        d = dirs[i, j]
        if d == DIAG:
            i, j = i - 1, j - 1
        elif d == UP:
            i -= 1
        else:
            j -= 1
        rev.append((i, j))
    return np.asarray(rev[::-1], dtype=np.int32)


def align_ref(a: np.ndarray, b: np.ndarray):
    """One pair: (cost fp32 scalar, path [L,2] int32)."""
    c = pair_cost(a, b)
    D, dirs = dtw_accumulate(c)
    return np.float32(D[-1, -1]), dtw_backtrack(dirs)


def phase_cost(c: np.ndarray, la: np.ndarray, lb: np.ndarray, penalty: float) -> np.ndarray:
    """Phase-conditioned cost: c'[i,j] = c[i,j] + (la[i] != lb[j] ? penalty : 0),
    one rounded fp32 add on top of pair_cost.  la [Ta], lb [Tb] integer phase labels."""
    c = np.asarray(c, dtype=np.float32)
    pen = np.where(np.asarray(la)[:, None] != np.asarray(lb)[None, :], np.float32(penalty), np.float32(0))
    with np.errstate(invalid="ignore", over="ignore"):
        return (c + pen.astype(np.float32)).astype(np.float32)


def align_phase_ref(a: np.ndarray, b: np.ndarray, la: np.ndarray, lb: np.ndarray, penalty: float):
    """One pair with phase labels: (cost fp32 scalar, path [L,2] int32)."""
    c = phase_cost(pair_cost(a, b), la, lb, penalty)
    D, dirs = dtw_accumulate(c)
    return np.float32(D[-1, -1]), dtw_backtrack(dirs)


def dtw_bruteforce(c: np.ndarray) -> float:
    """Minimum path cost by exhaustive enumeration (tiny Ta,Tb only); float64 sums,
    used by the oracle's own known-answer tests."""
    c = np.asarray(c, dtype=np.float64)
    Ta, Tb = c.shape
    best = [np.inf]

    def walk(i, j, acc):
        acc += c[i, j]
        if i == Ta - 1 and j == Tb - 1:
            best[0] = min(best[0], acc)
            return
        if i + 1 < Ta and j + 1 < Tb:
            walk(i + 1, j + 1, acc)
        if i + 1 < Ta:
            walk(i + 1, j, acc)
        if j + 1 < Tb:
            walk(i, j + 1, acc)

    walk(0, 0, 0.0)
    return best[0]


def compare_ref(a: np.ndarray, b: np.ndarray, path: np.ndarray) -> np.ndarray:
    """"Compare 2 skeleton" (README.md:50-52): per aligned step,
    per joint 2-D distance between the student and reference frames.
    a [Ta,V,>=2], b [Tb,V,>=2], path [L,2] -> [L,V] fp32 (same arithmetic as pair_cost)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ia, ib = path[:, 0], path[:, 1]
    dx = a[ia, :, 0] - b[ib, :, 0]
    dy = a[ia, :, 1] - b[ib, :, 1]
    s = dx * dx
    s = s + dy * dy
    return np.sqrt(s).astype(np.float32)


def synth_swings(N: int, Ta: int, Tb: int, V: int = 17, C: int = 2, seed: int = 7):
    """Synthetic swing pairs of SURVEY 8d: random-walk joints (cumsum of N(0,0.05)
    steps from an N(0,1) start) so warping paths are non-trivial."""
    rng = np.random.default_rng(seed)

    def walk(T):
        x0 = rng.standard_normal((N, 1, V, C)).astype(np.float32)
        steps = (rng.standard_normal((N, T, V, C)) * 0.05).astype(np.float32)
        return np.ascontiguousarray(x0 + np.cumsum(steps, axis=1, dtype=np.float32), dtype=np.float32)

    a = walk(Ta)
    b = walk(Tb)
    # the reference swing starts near the student's first pose, then drifts on its own
    b = b - b[:, :1] + a[:, :1] + (rng.standard_normal((N, 1, V, C)) * 0.1).astype(np.float32)
    return a, np.ascontiguousarray(b, dtype=np.float32)
