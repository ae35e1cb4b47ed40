"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's moving-horizon warm start for FullDiscretizationGridBase
(optimal_control/src/structured_ocp/discretization_grids/full_discretization_grid_base.cpp):
  findNearestState   :285-318   nearest state to the new measurement among the first min(N-2, 20) states, l2 norm, stop at the
                                first non-improving one; 0 if the start did not move (|dist| < 1e-12)
  warmStartShifting  :230-283   shift states/controls forward by num_shift, linear extrapolation of the tail states
                                x[idx] = x[idx-2] + 2 (x[idx-1] - x[idx-2]), controls repeated
  update             :95-107    x_seq.front() = x0 (measured), fixed goal components = xref
Pinned against the compiled reference by tests/golden/warm_start_shift.npz (tests/test_warm_start.py)."""
import numpy as np


def _norm(v):
    # Eigen's vectorised squaredNorm: SSE2 packets of two, two accumulators, scalar tail (Eigen/src/Core/Redux.h)
    t = v * v
    n = len(t)
    aligned, aligned2 = (n // 2) * 2, (n // 4) * 4
    if aligned == 0:
        return np.sqrt(t[0])
    p0 = t[0:2].copy()
    if aligned > 2:
        p1 = t[2:4].copy()
        for i in range(4, aligned2, 4):
            p0 += t[i:i + 2]
            p1 += t[i + 2:i + 4]
        p0 += p1
        if aligned > aligned2:
            p0 += t[aligned2:aligned2 + 2]
    res = p0[0] + p0[1]
    for i in range(aligned, n):
        res += t[i]
    return np.sqrt(res)


def find_nearest_state(x_seq, x0_new):
    """x_seq [N, nx] including the (old) start state and xf"""
    N = len(x_seq)
    first = _norm(x0_new - x_seq[0])
    if abs(first) < 1e-12:
        return 0
    lookahead = min((N - 1) - 1, 20)
    cache, nearest = first, 0
    for i in range(1, lookahead + 1):
        d = _norm(x0_new - x_seq[i])
        if d < cache:
            cache, nearest = d, i
        else:
            break
    return nearest


def warm_start_shift(x_seq, u_seq, x0_new):
    """x_seq [N, nx] (x_seq[N-1] = xf), u_seq [N-1, nu] -> shifted copies and num_shift; x_seq[0] is left to the caller (update :101)"""
    x, u = x_seq.copy(), u_seq.copy()
    N = len(x)
    s = find_nearest_state(x, x0_new)
    if s <= 0 or s > N - 2:
        return x, u, max(s, 0) if s <= N - 2 else s
    for i in range(N - s):
        idx = i + s
        if idx == N - 1:
            x[i] = x[N - 1]
        else:
            x[i] = x[idx]
            u[i] = u[idx]
    idx = N - s
    for i in range(s):
        x[idx] = x[idx - 2] + 2.0 * (x[idx - 1] - x[idx - 2])
        u[idx - 1] = u[idx - 2]
        idx += 1
    return x, u, s
